"""Secondary measurements for the other BASELINE configs (not the headline bench line):
  config 3  Mfcc(num_ceps=13, num_mel_bins=23) on 10 s cuts
  config 4  mixed 2-30 s cuts in duration-bucketed batches, padded (B, T_max, 80) output with LOG_EPSILON
            fill + frame lengths — the tensors OnTheFlyFeatures hands to K2SpeechRecognitionDataset
            (input_strategies.py:441-462), produced by one launch instead of extract_batch + collate_matrices
  staging   int16 PCM input (half the H2D bytes) through the same C-ABI host call
  kinds     Spectrogram / LogSpectrogram rows (257 floats per frame: write-heavy)
Each line: device-resident h/s (CUDA events) and host-to-host h/s (C-ABI extract_host, pinned buffers)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import lhotse_b200 as lb
from lhotse_b200.engine import OUT_PACKED, OUT_PADDED, Engine

SR = 16000
dev = torch.device("cuda", 0)


def time_device(eng, x, lens, offs, out_mode=OUT_PACKED, pad=0.0, reps=10):
    meta, tot = eng.plan_batch(lens, offs, out_mode=out_mode)
    meta_dev = torch.from_numpy(meta).to(dev)
    shape = (len(lens), tot.max_frames, eng.feature_dim) if out_mode == OUT_PADDED else (tot.total_rows, eng.feature_dim)
    out = torch.empty(shape, device=dev)
    for _ in range(3):
        eng.extract_device(x, lens, offs, out_mode=out_mode, pad_value=pad, out=out, meta_dev=meta_dev, totals=tot)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.extract_device(x, lens, offs, out_mode=out_mode, pad_value=pad, out=out, meta_dev=meta_dev, totals=tot)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3, out, tot


def time_host(eng, hx, lens, out_mode=OUT_PACKED, pad=0.0, reps=4):
    out, _ = eng.extract_host(hx, lens, out_mode=out_mode, pad_value=pad)
    hout = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
    eng.extract_host(hx, lens, out_mode=out_mode, pad_value=pad, out=hout)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.extract_host(hx, lens, out_mode=out_mode, pad_value=pad, out=hout)
    return (time.perf_counter() - t0) / reps


def report(name, eng, hours, t_dev, t_host, extra=None):
    rec = {"config": name, "kernel": eng.kernel, "device_h_per_s": hours / t_dev, "device_ms": t_dev * 1e3,
           "host_to_host_h_per_s": hours / t_host if t_host else None}
    rec.update(extra or {})
    print(json.dumps(rec), flush=True)


def main():
    torch.manual_seed(0)
    B, n = 1024, 160000
    x = 0.1 * torch.randn(B * n, device=dev)
    lens, offs = [n] * B, [i * n for i in range(B)]
    hx = torch.empty(B * n, dtype=torch.float32, pin_memory=True)
    hx.copy_(x)
    hours = B * n / SR / 3600

    for name, plan in (
        ("fbank80 10s (headline geometry)", lb.build_plan("fbank", lb.B200FbankConfig())),
        ("config3 mfcc13/23 10s", lb.build_plan("mfcc", lb.B200MfccConfig(num_ceps=13, num_mel_bins=23))),
        ("spectrogram257 10s", lb.build_plan("spectrogram", lb.B200SpectrogramConfig())),
        ("log-spectrogram257 10s", lb.build_plan("log-spectrogram", lb.B200LogSpectrogramConfig())),
    ):
        eng = Engine(plan, device=dev)
        t_dev, _, tot = time_device(eng, x, lens, offs)
        t_host = time_host(eng, hx, lens)
        report(name, eng, hours, t_dev, t_host, {"bytes_per_frame": 640 + 4 * eng.feature_dim,
                                                 "algorithmic_GBps": tot.total_rows * (640 + 4 * eng.feature_dim) / t_dev / 1e9})

    # plans outside the N=512 fast path run on the generic kernel (mixed-radix Stockham in shared memory)
    for name, cfg, sr in (
        ("fbank80 16k N=400 (round_to_power_of_two=False; fast400, prime-factor 8x25)", lb.B200FbankConfig(round_to_power_of_two=False), 16000),
        ("fbank80 16k N=400 forced generic (radices 4,2,5,5)", lb.B200FbankConfig(round_to_power_of_two=False, kernel="generic"), 16000),
        ("fbank40 8k N=256", lb.B200FbankConfig(sampling_rate=8000, num_filters=40), 8000),
        ("fbank80 24k N=1024 (fast1024)", lb.B200FbankConfig(sampling_rate=24000), 24000),
        ("fbank80 22.05k N=1024 (fast1024)", lb.B200FbankConfig(sampling_rate=22050), 22050),
        ("fbank80 24k N=1024 forced generic", lb.B200FbankConfig(sampling_rate=24000, kernel="generic"), 24000),
        ("fbank80 24k 50ms N=2048 (fast2048)", lb.B200FbankConfig(sampling_rate=24000, frame_length=0.05), 24000),
        ("fbank80 44.1k N=2048 (fast2048)", lb.B200FbankConfig(sampling_rate=44100), 44100),
        ("fbank80 48k N=2048 (fast2048)", lb.B200FbankConfig(sampling_rate=48000), 48000),
        ("fbank80 24k 50ms N=2048 forced generic", lb.B200FbankConfig(sampling_rate=24000, frame_length=0.05, kernel="generic"), 24000),
        ("fbank80 16k N=512 forced generic", lb.B200FbankConfig(kernel="generic"), 16000),
    ):
        eng = Engine(lb.build_plan("fbank", cfg), device=dev, kernel=getattr(cfg, "kernel", "auto"))
        nn = 10 * sr
        Bg = 256 if sr <= 24000 else 128
        xg = x[: Bg * nn]
        lg, og = [nn] * Bg, [i * nn for i in range(Bg)]
        t_dev, _, tot = time_device(eng, xg, lg, og, reps=5)
        report(name, eng, Bg * 10 / 3600, t_dev, None, {"frames": int(tot.total_rows)})

    # int16 staging: same kernel, half the input bytes over PCIe and HBM
    eng = Engine(lb.build_plan("fbank", lb.B200FbankConfig()), device=dev)
    xi = (x * 32767).clamp(-32768, 32767).to(torch.int16)
    hxi = torch.empty(B * n, dtype=torch.int16, pin_memory=True)
    hxi.copy_(xi)
    t_dev, _, _ = time_device(eng, xi, lens, offs)
    t_host = time_host(eng, hxi, lens)
    report("fbank80 10s, int16 PCM staging", eng, hours, t_dev, t_host)

    # config 4 stand-in: 2-30 s cuts, batches of ~600 s bucketed by duration, padded collation output
    rs = np.random.RandomState(0)
    durs = np.sort(rs.uniform(2.0, 30.0, size=4096))
    batches, cur, acc = [], [], 0.0
    for d in durs:  # neighbours in the sorted list = one duration bucket
        if acc + d > 600.0 and cur:
            batches.append(cur); cur, acc = [], 0.0
        cur.append(int(d * SR)); acc += d
    if cur:
        batches.append(cur)
    tot_hours = sum(sum(b) for b in batches) / SR / 3600
    big = 0.1 * torch.randn(max(sum(b) + 4 * len(b) for b in batches), device=dev)
    t_all, pad_frac = 0.0, []
    for blens in batches:
        offs_b, curo = [], 0
        for m in blens:
            curo = (curo + 3) // 4 * 4
            offs_b.append(curo); curo += m
        t, out, tot = time_device(eng, big, blens, offs_b, out_mode=OUT_PADDED, pad=lb.LOG_EPSILON, reps=3)
        t_all += t
        pad_frac.append(1.0 - tot.total_rows / (len(blens) * tot.max_frames))
    report("config4 mixed 2-30s, bucketed ~600s batches, padded (B,Tmax,80)+LOG_EPSILON", eng, tot_hours, t_all, None,
           {"batches": len(batches), "mean_pad_fraction": float(np.mean(pad_frac))})


if __name__ == "__main__":
    main()
