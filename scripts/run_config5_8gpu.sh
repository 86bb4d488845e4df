#!/bin/bash
# CutSet-level store of BASELINE configs[4] on all GPUs of the box: one rank per GPU, then two ranks per GPU (gloo collectives).
out=${1:-gpurun_out/r2_config5_8gpu_v2.jsonl}
: > "$out"
n=$(nvidia-smi -L | wc -l)
for ranks in $n $((2 * n)); do
  timeout 300 python -m torch.distributed.run --nproc-per-node $ranks --master-addr 127.0.0.1 --master-port $((29600 + ranks)) \
      scripts/bench_config5.py --hours-per-rank 10 --num-workers 4 2>&1 | grep "^{" | head -1 | tee -a "$out" | cut -c1-260
done
