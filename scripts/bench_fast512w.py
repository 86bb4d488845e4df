"""Headline workload (Fbank-80, 16 kHz, 2048 x 10 s cuts, device-resident) on the half-warp-per-frame kernel (fast512.cuh) and on the
warp-per-frame kernel (fast512w.cuh, B200FEAT_FAST_VARIANT=3) in its launch shapes; outputs compared with each other."""
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
    B, n = int(os.environ.get("F5W_B", "2048")), 160000
    torch.manual_seed(0)
    x = 0.1 * torch.randn(B * n, device=dev)
    lens, offs = [n] * B, [i * n for i in range(B)]
    ref = None
    for kind, cfg in (("fbank", lb.B200FbankConfig()), ("mfcc", lb.B200MfccConfig(num_ceps=13, num_mel_bins=23))):
        for variant, shape in (("0", ""), ("3", "0"), ("3", "1"), ("3", "2"), ("3", "3")):
            os.environ["B200FEAT_FAST_VARIANT"] = variant
            os.environ["B200FEAT_FAST512W_SHAPE"] = shape or "0"
            eng = Engine(lb.build_plan(kind, cfg), device=dev, kernel="fast")
            t, out, tot = time_device(eng, x, lens, offs, reps=20)
            chk = out[:: max(1, out.shape[0] // 4096)].double().cpu().numpy()
            if variant == "0":
                ref = chk
            print(json.dumps({"feature": kind, "kernel": "fast512" if variant == "0" else "fast512w", "shape": shape, "h_per_s": B * 10 / 3600 / t,
                              "ms": t * 1e3, "max_abs_diff_vs_fast512": float(np.abs(chk - ref).max())}), flush=True)


if __name__ == "__main__":
    main()
