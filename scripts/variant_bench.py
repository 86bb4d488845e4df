"""Developer tool: times the headline workload (Fbank-80, 2048 x 10 s cuts, device-resident) once per library variant in
build_variants/ (plus the in-tree library), each in its own process, and checks every variant's output against the
in-tree library's.  Run on the GPU box:  python scripts/variant_bench.py [B] [reps]"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
import lhotse_b200 as lb
from lhotse_b200.engine import Engine
B, reps, feature = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
n = 160000
dev = torch.device("cuda", 0)
cfgs = {"fbank": lb.B200FbankConfig(), "fbank400": lb.B200FbankConfig(round_to_power_of_two=False), "mfcc": lb.B200MfccConfig(num_ceps=13, num_filters=23), "spectrogram": lb.B200SpectrogramConfig()}
eng = Engine(lb.build_plan("fbank" if feature.startswith("fbank") else feature, cfgs[feature]), device=dev, kernel="fast")
torch.manual_seed(0)
x = 0.1 * torch.randn(B * n, device=dev)
lens, offs = [n] * B, [i * n for i in range(B)]
meta, tot = eng.plan_batch(lens, offs)
meta_dev = torch.from_numpy(meta).to(dev)
out = torch.empty((tot.total_rows, eng.feature_dim), device=dev)
for _ in range(3):
    eng.extract_device(x, lens, offs, out=out, meta_dev=meta_dev, totals=tot)
torch.cuda.synchronize()
best = 1e9
for rnd in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.extract_device(x, lens, offs, out=out, meta_dev=meta_dev, totals=tot)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / reps)
chk = out[:: max(1, out.shape[0] // 4096)].double().cpu().numpy()
np.save(sys.argv[4], chk)
print(json.dumps({"ms": best, "h_per_s": B * n / 16000 / 3600 / (best / 1e3)}))
''' % ROOT


def main():
    B = sys.argv[1] if len(sys.argv) > 1 else "2048"
    reps = sys.argv[2] if len(sys.argv) > 2 else "20"
    feature = sys.argv[3] if len(sys.argv) > 3 else "fbank"
    libs = [("in-tree", None)] + [(os.path.basename(p)[len("libb200feat_"):-3], p)
                                  for p in sorted(glob.glob(os.path.join(ROOT, "build_variants", "libb200feat_*.so")))]
    import numpy as np
    ref = None
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for name, path in libs:
        env = dict(os.environ)
        if path:
            env["B200FEAT_LIBRARY"] = path
        dump = os.path.join(ROOT, "gpurun_out", f"variant_{name}.npy")
        res = subprocess.run([sys.executable, "-c", CHILD, B, reps, feature, dump], env=env, capture_output=True, text=True)
        if res.returncode != 0:
            print(name, "FAILED", res.stderr[-500:])
            continue
        rec = json.loads(res.stdout.strip().splitlines()[-1])
        got = np.load(dump)
        os.remove(dump)
        if ref is None:
            ref = got
        rec["max_abs_diff_vs_in_tree"] = float(np.abs(got - ref).max())
        print(f"{name:28s} {rec['h_per_s']:8.1f} h/s  {rec['ms']:.4f} ms  maxdiff {rec['max_abs_diff_vs_in_tree']:.2e}", flush=True)


if __name__ == "__main__":
    main()
