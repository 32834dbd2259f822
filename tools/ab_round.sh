#!/bin/sh
# GPU tests + same-box A/B of variants/libdcsim_base.so against the in-tree build + per-option device time
{
(timeout 600 python -m pytest tests -m gpu -x -q) 2>&1 | tail -1
sh tools/ab_bench.sh variants/libdcsim_base.so distributed_cluster_gpus_b200/csrc/libdcsim_b200.so
timeout 300 python tools/time_options.py 2>&1 | tail -5
} > gpurun_out/ab_round.log 2>&1
cat gpurun_out/ab_round.log
