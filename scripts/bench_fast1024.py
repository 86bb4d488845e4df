"""N = 1024 geometries (24 kHz and 22.05 kHz with 25 ms frames, 16 kHz / 64 ms): device-resident h/s of the fast1024 kernel in its launch
shapes (B200FEAT_FAST1024_VARIANT) against the generic kernel, 10 s cuts."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import lhotse_b200 as lb
from lhotse_b200.engine import Engine
from scripts.bench_configs import time_device

dev = torch.device("cuda", 0)


def main():
    torch.manual_seed(0)
    geos = (("24k/25ms L=600 S=240", dict(sampling_rate=24000), 24000), ("22.05k/25ms L=551 S=220", dict(sampling_rate=22050), 22050),
            ("16k/64ms L=N=1024 S=256", dict(sampling_rate=16000, frame_length=0.064, frame_shift=0.016), 16000))
    for name, cfg, sr in geos:
        B, nn = 256, 10 * sr
        x = 0.1 * torch.randn(B * nn, device=dev)
        lens, offs = [nn] * B, [i * nn for i in range(B)]
        ref = None
        runs = [("generic", "", "")] + [("fast", v, "") for v in os.environ.get("F1K_VARIANTS", "0,1,2,3").split(",")]
        for kernel, variant, old in runs:
            os.environ["B200FEAT_FAST1024_VARIANT"] = variant or "0"
            os.environ.pop("B200FEAT_FAST1024_OLD", None)
            if old:
                os.environ["B200FEAT_FAST1024_OLD"] = "1"
            eng = Engine(lb.build_plan("fbank", lb.B200FbankConfig(kernel=kernel, **cfg)), device=dev, kernel=kernel)
            t, out, tot = time_device(eng, x, lens, offs, reps=10)
            chk = out[:: max(1, out.shape[0] // 4096)].double().cpu().numpy()
            if ref is None:
                ref = chk
            print(json.dumps({"geometry": name, "kernel": eng.kernel + ("-round1" if old else ""), "variant": variant, "h_per_s": B * 10 / 3600 / t,
                              "ms": t * 1e3, "frames": int(tot.total_rows), "max_abs_diff_vs_generic": float(np.abs(chk - ref).max())}), flush=True)


if __name__ == "__main__":
    main()
