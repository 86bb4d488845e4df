"""Secondary measurement: LibrosaFbank geometry (22.05 kHz, fft 1024, hop 256, 80 Slaney mels) on 10 s cuts, device-resident."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lhotse_b200 as lb
from lhotse_b200.engine import Engine

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, n, sr = 512, 220500, 22050
x = 0.1 * torch.randn(B * n, device=dev)
lens, offs = [n] * B, [i * n for i in range(B)]
for kernel in ("auto", "generic"):
    Bk = B if kernel == "auto" else 64
    eng = Engine(lb.build_plan("librosa-fbank", lb.B200LibrosaFbankConfig()), device=dev, kernel=kernel)
    meta, tot = eng.plan_batch(lens[:Bk], offs[:Bk])
    meta_dev = torch.from_numpy(meta).to(dev)
    out = torch.empty(int(tot.out_floats), device=dev)
    for _ in range(3):
        eng.extract_device(x, lens[:Bk], offs[:Bk], out=out, meta_dev=meta_dev, totals=tot)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        eng.extract_device(x, lens[:Bk], offs[:Bk], out=out, meta_dev=meta_dev, totals=tot)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 10 / 1e3
    print(json.dumps({"config": "librosa-fbank 22.05k fft1024 hop256 80 mels, 10 s cuts", "kernel": eng.kernel, "cuts": Bk,
                      "device_ms": t * 1e3, "device_h_per_s": Bk * n / sr / 3600 / t, "frames": int(tot.total_rows)}), flush=True)
