"""Host<->device copy rates on this box, to judge how close the e2e path (H2D of float32 samples overlapped with the
kernel and the D2H of features) is to the link: pinned H2D alone, D2H alone, both directions at once, by chunk size."""
import time
import torch

dev = torch.device("cuda", 0)
nbytes = 640 << 20
h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(nbytes // 2, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d_out = torch.empty(nbytes // 2, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(chunk, both, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for o in range(0, nbytes, chunk):
            with torch.cuda.stream(s1):
                d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    h_out[o // 2:(o + chunk) // 2].copy_(d_out[o // 2:(o + chunk) // 2], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for chunk in (8 << 20, 32 << 20, 128 << 20, nbytes):
    t = run(chunk, False)
    tb = run(chunk, True)
    print(f"chunk {chunk >> 20:4d} MiB: H2D alone {nbytes / t / 1e9:6.1f} GB/s | H2D+D2H(half) together: H2D {nbytes / tb / 1e9:6.1f} GB/s "
          f"(+ D2H {nbytes / 2 / tb / 1e9:5.1f} GB/s)")
