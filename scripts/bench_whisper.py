"""Secondary measurement: Whisper log-mel front end, 80 filters, 10 s cuts.
Prints one JSON line per arm: device-resident h/s (CUDA events around the launch pair: fused N = 400 kernel + normalise
pass), host-to-host h/s through the C-ABI host call, and — for scale — the reference's own torch op chain
(whisper_fbank.py:16-84: torch.stft + matmul + log10 ...) run on CUDA tensors on the same GPU, one cut per call as
`WhisperFbank(device="cuda").extract` does."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import lhotse_b200 as lb
from lhotse_b200.engine import Engine

SR = 16000
dev = torch.device("cuda", 0)


def torch_chain(audio, filters, window):
    stft = torch.stft(audio, 400, 160, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    v = torch.clamp(filters @ mag, min=1e-10).log10()
    v = torch.maximum(v, v.max() - 8.0)
    return ((v + 4.0) / 4.0).transpose(0, 1)


def main():
    torch.manual_seed(0)
    B, n = int(os.environ.get("CUTS", 1024)), 160000
    x = 0.1 * torch.randn(B * n, device=dev)
    lens, offs = [n] * B, [i * n for i in range(B)]
    hours = B * n / SR / 3600
    for kernel in ("auto", "generic"):
        Bk = B if kernel == "auto" else min(B, 256)
        eng = Engine(lb.build_plan("whisper-fbank", lb.B200WhisperFbankConfig()), device=dev, kernel=kernel)
        meta, tot = eng.plan_batch(lens[:Bk], offs[:Bk])
        meta_dev = torch.from_numpy(meta).to(dev)
        out = torch.empty(int(tot.out_floats), device=dev)
        for _ in range(3):
            eng.extract_device(x, lens[:Bk], offs[:Bk], out=out, meta_dev=meta_dev, totals=tot)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a.record()
        for _ in range(reps):
            eng.extract_device(x, lens[:Bk], offs[:Bk], out=out, meta_dev=meta_dev, totals=tot)
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) / reps / 1e3
        rec = {"config": "whisper-fbank80 10s", "kernel": eng.kernel, "cuts": Bk, "device_ms": t * 1e3,
               "device_h_per_s": Bk * n / SR / 3600 / t, "launches_per_batch": 2,
               "algorithmic_GBps": tot.total_rows * (640 + 320) / t / 1e9}
        if kernel == "auto":
            hx = torch.empty(B * n, dtype=torch.float32, pin_memory=True)
            hx.copy_(x)
            o, _ = eng.extract_host(hx, lens)
            hout = torch.empty(o.shape, dtype=torch.float32, pin_memory=True)
            eng.extract_host(hx, lens, out=hout)
            t0 = time.perf_counter()
            for _ in range(4):
                eng.extract_host(hx, lens, out=hout)
            rec["host_to_host_h_per_s"] = hours / ((time.perf_counter() - t0) / 4)
        print(json.dumps(rec), flush=True)

    # the reference's op chain on the same GPU (per-cut calls, as WhisperFbank.extract works)
    window = torch.hann_window(400, device=dev)
    filters = torch.from_numpy(lb.build_plan("whisper-fbank", lb.B200WhisperFbankConfig()).mel_bank.T.copy()).to(dev)
    cuts = [x[i * n:(i + 1) * n] for i in range(64)]
    for c in cuts[:8]:
        torch_chain(c, filters, window)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c in cuts:
        torch_chain(c, filters, window)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(json.dumps({"config": "whisper-fbank80 10s", "kernel": "torch op chain on cuda (reference's GPU path), per cut",
                      "cuts": 64, "device_h_per_s": 64 * n / SR / 3600 / t}), flush=True)


if __name__ == "__main__":
    main()
