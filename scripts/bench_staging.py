"""Secondary measurements for the staging rows:
  1. what an odd sample offset costs the kernel (cuts staged back to back after an odd-length cut leave the vector-load
     path) — the reason `extract_batch` / `b200feat_extract_host_at` stage every cut on a 4-element boundary;
  2. file -> features: 16-bit PCM WAV files through the pinned int16 ring (`lhotse_b200.pcm_staging`) against the float
     route (decode to float32 on the host, stage, H2D 4 bytes per sample), both ending in the padded device tensor."""
import json
import os
import sys
import tempfile
import time
import wave

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import lhotse_b200 as lb
from lhotse_b200.pcm_staging import PcmRequest, PcmStagingRing

dev = torch.device("cuda", 0)
SR = 16000


def time_device(eng, x, lens, offs, reps=10):
    meta, tot = eng.plan_batch(lens, offs)
    meta_dev = torch.from_numpy(meta).to(dev)
    out = torch.empty(int(tot.out_floats), device=dev)
    for _ in range(3):
        eng.extract_device(x, lens, offs, out=out, meta_dev=meta_dev, totals=tot)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.extract_device(x, lens, offs, out=out, meta_dev=meta_dev, totals=tot)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3


def main():
    torch.manual_seed(0)
    ext = lb.B200Fbank()
    eng = ext.engine
    B, n = 1024, 159999  # odd length
    x = 0.1 * torch.randn(B * (n + 1) + 8, device=dev)
    lens = [n] * B
    hours = B * n / SR / 3600
    for name, offs in (("back to back (every second cut on an odd offset)", [i * n for i in range(B)]),
                       ("4-aligned offsets", [i * (n + 1) for i in range(B)])):
        t = time_device(eng, x, lens, offs)
        print(json.dumps({"what": "device-resident Fbank-80, 1024 cuts of 159999 samples", "staging": name,
                          "ms": t * 1e3, "h_per_s": hours / t}), flush=True)

    # ---- file -> padded device features
    rs = np.random.RandomState(0)
    nf, ns = 256, 160000
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    paths = []
    for i in range(nf):
        p = os.path.join(d, f"{i}.wav")
        with wave.open(p, "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(SR)
            w.writeframes(np.clip(rs.randn(ns) * 3000, -32768, 32767).astype("<i2").tobytes())
        paths.append(p)
    fh = nf * ns / SR / 3600
    ring = PcmStagingRing()
    reqs = [PcmRequest(p, 0, ns) for p in paths]

    def pcm_route():
        staged, lens_, offs_, sr = ring.stage(reqs)
        f, _ = ext.extract_staged_padded(staged, lens_, offs_, sr)
        return f

    def float_route():  # what a libsndfile-style decoder hands the extractor: float32 arrays, then extract_batch_padded
        waves = []
        for p in paths:
            with wave.open(p) as w:
                waves.append(torch.from_numpy(np.frombuffer(w.readframes(ns), dtype="<i2").astype(np.float32) / 32768.0))
        f, _ = ext.extract_batch_padded(waves, SR)
        return f

    for name, fn in (("pcm16 ring (file -> pinned int16 -> GPU)", pcm_route), ("float route (decode to float32 -> stage -> GPU)", float_route)):
        a = fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / 4
        print(json.dumps({"what": "256 WAV files x 10 s (tmpfs) -> padded (B, T, 80) on the device, 1 host thread",
                          "route": name, "ms": t * 1e3, "h_per_s": fh / t}), flush=True)
    assert torch.equal(pcm_route(), float_route())
    for p in paths:
        os.remove(p)
    os.rmdir(d)


if __name__ == "__main__":
    main()
