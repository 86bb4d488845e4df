import os, sys
sys.path.insert(0, "/root/repo")
os.environ["B200FEAT_FAST_VARIANT"] = "3"; os.environ["B200FEAT_FAST512W_SHAPE"] = os.environ.get("SHAPE", "3")
import torch, lhotse_b200 as lb
from lhotse_b200.engine import Engine
from scripts.bench_configs import time_device
dev = torch.device("cuda", 0)
B, n = 512, 160000
x = 0.1 * torch.randn(B * n, device=dev)
eng = Engine(lb.build_plan("fbank", lb.B200FbankConfig()), device=dev, kernel="fast")
t, _, _ = time_device(eng, x, [n] * B, [i * n for i in range(B)], reps=1)
print(t)
