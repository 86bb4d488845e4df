"""Why is the int16 host route slower than the float32 one through extract_batch?  Times the pieces (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lhotse_b200 as lb

B, n = 1024, 160000
ext = lb.B200Fbank(lb.B200FbankConfig(device="cuda:0"))
eng = ext.engine
x = (0.1 * torch.randn(B, n))
hf = torch.empty((B, n), dtype=torch.float32, pin_memory=True); hf.copy_(x)
hi = torch.empty((B, n), dtype=torch.int16, pin_memory=True); hi.copy_((x * 32767).clamp_(-32768, 32767).to(torch.int16))
lens = [n] * B

def t(fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3

out = torch.empty((B * 1000, 80), dtype=torch.float32, pin_memory=True).numpy()
print("extract_host f32 flat, preallocated out : %.2f ms" % t(lambda: eng.extract_host(hf.view(-1).numpy(), lens, out=out)))
print("extract_host i16 flat, preallocated out : %.2f ms" % t(lambda: eng.extract_host(hi.view(-1).numpy(), lens, out=out)))
print("extract_host f32 flat, fresh out        : %.2f ms" % t(lambda: eng.extract_host(hf.view(-1).numpy(), lens)))
print("extract_host i16 flat, fresh out        : %.2f ms" % t(lambda: eng.extract_host(hi.view(-1).numpy(), lens)))
print("extract_batch f32 (B, n)                : %.2f ms" % t(lambda: ext.extract_batch(hf.numpy(), 16000)))
print("extract_batch i16 (B, n)                : %.2f ms" % t(lambda: ext.extract_batch(hi.numpy(), 16000)))

# --- replicate bench.py's sequence step by step
from lhotse_b200 import dist as lbd
xg = x.cuda()
h2 = torch.empty((B, n), dtype=torch.int16, pin_memory=True)
h2.copy_((xg.view(B, n) * 32767.0).clamp_(-32768, 32767).to(torch.int16))
print("i16 pinned filled from a CUDA tensor     : %.2f ms" % t(lambda: ext.extract_batch(h2.numpy(), 16000)))
ext2 = lb.B200Fbank(lb.B200FbankConfig(device="cuda:0"))
print("same, fresh extractor                    : %.2f ms" % t(lambda: ext2.extract_batch(h2.numpy(), 16000)))
node = lbd.bind_host_to_gpu_numa(0)
print("after bind_host_to_gpu_numa -> node", node)
print("i16 (old pinned buffers) after binding   : %.2f ms" % t(lambda: ext2.extract_batch(h2.numpy(), 16000)))
h3 = torch.empty((B, n), dtype=torch.int16, pin_memory=True); h3.copy_(h2)
ext3 = lb.B200Fbank(lb.B200FbankConfig(device="cuda:0"))
print("i16 new pinned buffer + extractor after binding: %.2f ms" % t(lambda: ext3.extract_batch(h3.numpy(), 16000)))
print("f32 after binding                        : %.2f ms" % t(lambda: ext3.extract_batch(hf.numpy(), 16000)))
import time as _t
a16 = h3.numpy()
for _ in range(2): ext3.extract_batch(a16, 16000)
t0 = _t.perf_counter(); k = 0
while k < 6:
    f = ext3.extract_batch(a16, 16000); k += 1
print("bench-style loop: %.2f ms per call" % ((_t.perf_counter() - t0) / 6 * 1e3))
