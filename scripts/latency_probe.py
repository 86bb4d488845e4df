"""Per-call latency of the public API for single cuts (what CutSet.compute_and_store_features pays per cut)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import lhotse_b200 as lb
ext = lb.B200Fbank()
rs = np.random.RandomState(0)
for secs in (1, 10, 30):
    x = (0.1 * rs.randn(secs * 16000)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    for name, fn in (("extract(numpy)", lambda: ext.extract(x, 16000)),
                     ("extract(cuda tensor)", lambda: ext.extract(xt, 16000)),
                     ("extract_batch([numpy]*8)", lambda: ext.extract_batch([x] * 8, 16000))):
        for _ in range(5): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(50): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
        print(f"{secs:3d}s {name:28s} {dt*1e6:9.1f} us/call")
import cProfile, pstats
x = (0.1 * rs.randn(160000)).astype(np.float32)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): ext.extract(x, 16000)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
