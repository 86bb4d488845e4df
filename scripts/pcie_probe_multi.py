#!/usr/bin/env python
"""Per-rank pinned H2D / D2H rates with 1, 2, 4, ... ranks copying AT THE SAME TIME (torchrun): the diagnosis behind the e2e
scaling of bench.py (VERDICT r1 'weak' #5: 1 -> 8 GPUs at 0.62 while the device-resident path scales at 0.996).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/pcie_probe_multi.py

For every active-rank count k in {1, 2, 4, 8 (<= world)} the first k ranks move 640 MiB host->device together with 320 MiB
device->host (the e2e traffic pattern of Fbank-80 on float32 samples) while the others idle; reports the per-rank and the
aggregate rates, with and without binding every rank to its GPU's NUMA node (B200FEAT_NO_NUMA_BIND=1 for the unbound run),
and where the pinned buffers live.  One JSON line on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from lhotse_b200 import dist as lbd  # noqa: E402


def main():
    rank, world, local = lbd.init_distributed()
    torch.cuda.set_device(local)
    node = lbd.bind_host_to_gpu_numa(local)
    dev = torch.device("cuda", local)
    nbytes = 640 << 20
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes // 2, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(nbytes // 2, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    chunk = 32 << 20

    def run(reps=4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for o in range(0, nbytes, chunk):
                with torch.cuda.stream(s1):
                    d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
                with torch.cuda.stream(s2):
                    h_out[o // 2:(o + chunk) // 2].copy_(d_out[o // 2:(o + chunk) // 2], non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    run(1)
    rows = []
    k = 1
    while k <= world:
        lbd.barrier()
        t = run() if rank < k else 0.0
        lbd.barrier()
        ts = [0.0] * world
        ts[rank] = t
        ts = lbd.all_reduce_stats(ts, "sum")
        act = [x for x in ts[:k]]
        rows.append({"active_ranks": k, "h2d_gbs_per_rank_min": nbytes / max(act) / 1e9, "h2d_gbs_per_rank_max": nbytes / min(act) / 1e9,
                     "h2d_gbs_aggregate": sum(nbytes / x for x in act) / 1e9, "d2h_gbs_aggregate": sum(nbytes / 2 / x for x in act) / 1e9})
        k *= 2
    nodes = lbd.all_reduce_stats([float(node if node is not None else -1) if r == rank else 0.0 for r in range(world)], "sum")
    if rank == 0:
        print(json.dumps({"world": world, "numa_bind": os.environ.get("B200FEAT_NO_NUMA_BIND") != "1", "gpu_numa_node_per_rank": nodes,
                          "host_cpus": os.cpu_count(), "pattern": "640 MiB H2D + 320 MiB D2H per rank, 32 MiB chunks, pinned", "rows": rows}), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
