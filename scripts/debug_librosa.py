import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lhotse_b200 as lb
from oracle import librosa_oracle as LO
g = np.load(os.path.join(ROOT, "tests/golden/golden_librosa_v1.npz"))
man = json.loads(bytes(g["manifest"]).decode())
for i in (9, 10, 11):
    c, x = man[i], g[f"x{i}"]
    cfg = c["cfg"]
    t = LO.extract(x, float64=True, **cfg)
    for k in ("fast", "generic"):
        got = lb.B200LibrosaFbank(lb.B200LibrosaFbankConfig(kernel=k, **cfg)).extract(x, cfg["sampling_rate"])
        d = np.abs(got - t)
        rows = np.where(d.max(axis=1) > 1e-3)[0]
        print(i, cfg["fft_size"], cfg["hop_size"], k, "max", d.max(), "bad rows", rows[:20], len(rows), "of", got.shape[0],
              "bad cols of first bad row", (np.where(d[rows[0]] > 1e-3)[0][:12] if len(rows) else None))
