"""`extract_batch(list of separately allocated numpy arrays)` (the C-side gather, b200feat_extract_host_ptrs) against the number of
gather threads (B200FEAT_STAGING_THREADS; one handle per setting).  1024 x 10 s cuts, float32."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import lhotse_b200 as lb

SR, B, n = 16000, 1024, 160000
rs = np.random.RandomState(0)
lst = [(0.1 * rs.randn(n)).astype(np.float32) for _ in range(B)]
for th in sys.argv[1:] or ["4", "8", "12", "16", "24"]:
    os.environ["B200FEAT_STAGING_THREADS"] = th
    ext = lb.B200Fbank()
    for _ in range(2):
        ext.extract_batch(lst, SR)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        out = ext.extract_batch(lst, SR)
    dt = (time.perf_counter() - t0) / 4
    print(json.dumps({"gather_threads": int(th), "h_per_s": B * n / SR / 3600 / dt, "ms": dt * 1e3, "gather_GBps": B * n * 4 / dt / 1e9}), flush=True)
    ext.engine.close()
