"""N = 2048 geometries (24 kHz / 50 ms, 44.1 kHz and 48 kHz / 25 ms): device-resident h/s of the fast2048 kernel in its launch
shapes (B200FEAT_FAST2048_VARIANT) against the generic kernel, on 10 s cuts.  `--one` runs a single launch (for ncu)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import lhotse_b200 as lb
from lhotse_b200.engine import Engine
from scripts.bench_configs import time_device

dev = torch.device("cuda", 0)


def main():
    one = "--one" in sys.argv
    torch.manual_seed(0)
    geos = (("24k/50ms L=1200 S=240", dict(sampling_rate=24000, frame_length=0.05), 24000),
            ("44.1k/25ms L=1102 S=441", dict(sampling_rate=44100), 44100),
            ("48k/25ms L=1200 S=480", dict(sampling_rate=48000), 48000),
            ("16k/128ms L=N=2048 S=512", dict(sampling_rate=16000, frame_length=0.128, frame_shift=0.032), 16000))
    for name, cfg, sr in geos[:1] if one else geos[:int(os.environ.get("F2K_GEOS", "4"))]:
        B, nn = (256 if sr <= 24000 else 128), 10 * sr
        x = 0.1 * torch.randn(B * nn, device=dev)
        lens, offs = [nn] * B, [i * nn for i in range(B)]
        for kernel, variant in (("fast", "0"),) if one else [("fast", v) for v in os.environ.get("F2K_VARIANTS", "0,1,2,3").split(",")] + [("generic", "")] * ("--no-generic" not in sys.argv):
            os.environ["B200FEAT_FAST2048_VARIANT"] = variant or "0"
            eng = Engine(lb.build_plan("fbank", lb.B200FbankConfig(kernel=kernel, **cfg)), device=dev, kernel=kernel)
            t, _, tot = time_device(eng, x, lens, offs, reps=1 if one else 10)
            print(json.dumps({"geometry": name, "kernel": eng.kernel, "variant": variant, "h_per_s": B * 10 / 3600 / t, "ms": t * 1e3,
                              "frames": int(tot.total_rows), "ns_per_frame": t * 1e9 / tot.total_rows}), flush=True)


if __name__ == "__main__":
    main()
