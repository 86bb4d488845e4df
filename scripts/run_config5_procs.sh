#!/bin/bash
# CutSet-level store (scripts/bench_config5.py) with 1, 2, 4, 8 processes sharing ONE GPU (dist.init_distributed maps local rank r to
# GPU r mod #GPUs and uses gloo for the job-level collectives): the job is bound by per-cut Python, not by the GPU.
out=${1:-gpurun_out/r2_config5_procs_per_gpu.jsonl}
: > "$out"
nproc
for n in 1 2 4 8; do
  timeout 200 python -m torch.distributed.run --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      scripts/bench_config5.py --hours-per-rank 5 --num-workers 4 2>&1 | grep "^{" | head -1 | sed "s/^{/{\"procs_on_one_gpu\": $n, /" | tee -a "$out" | cut -c1-230
done
