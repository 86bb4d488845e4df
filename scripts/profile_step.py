"""A short device-resident run of the headline hot path for `ncu` (one GPU, few launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lhotse_b200 as lb
from lhotse_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kernel = sys.argv[2] if len(sys.argv) > 2 else "auto"
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 6
n = 160000
dev = torch.device("cuda", 0)
geom = sys.argv[4] if len(sys.argv) > 4 else "n512"   # n512 | n400 (round_to_power_of_two=False) | whisper
if geom == "whisper":
    plan = lb.build_plan("whisper-fbank", lb.B200WhisperFbankConfig())
else:
    plan = lb.build_plan("fbank", lb.B200FbankConfig(round_to_power_of_two=(geom != "n400")))
eng = Engine(plan, device=dev, kernel=kernel)
torch.manual_seed(0)
x = 0.1 * torch.randn(B * n, device=dev)
lens, offs = [n] * B, [i * n for i in range(B)]
meta, tot = eng.plan_batch(lens, offs)
meta_dev = torch.from_numpy(meta).to(dev)
out = torch.empty(int(tot.out_floats), device=dev)
for _ in range(launches):
    eng.extract_device(x, lens, offs, out=out, meta_dev=meta_dev, totals=tot)
torch.cuda.synchronize()
print("kernel", eng.kernel, "frames", tot.total_rows, "finite", bool(torch.isfinite(out[: tot.total_rows * eng.feature_dim]).all()))
