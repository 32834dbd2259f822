#!/bin/sh
# final-build evidence: sanitizers (lean + full records), launch list, ncu full of both kernels, bench both arms, full parity
sh tools/gpu_check.sh 2>&1 | grep -v "^\." | tail -8
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v6.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e) > gpurun_out/r1_launches_v6.log 2>&1; echo launch-list $?
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcsim_advance -s 1 -c 1 -f -o gpurun_out/r1_prof_v6_advance python bench.py --steps 1 --warmup 1 --replicas 8192 --no-cpu-baseline --no-e2e) > gpurun_out/r1_ncu_full_v6a.log 2>&1; echo full-advance $?
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcsim_arrivals -s 1 -c 1 -f -o gpurun_out/r1_prof_v6_arrivals python bench.py --steps 1 --warmup 1 --replicas 8192 --no-cpu-baseline --no-e2e) > gpurun_out/r1_ncu_full_v6b.log 2>&1; echo full-arrivals $?
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1_bench_v6_reference.json 2> gpurun_out/r1_bench_v6_reference.err; echo ref-arm $?
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench_v6.json 2> gpurun_out/r1_bench_v6.err; echo bench $?
tail -c 600 gpurun_out/r1_bench_v6.json
timeout 900 python tools/full_parity.py > gpurun_out/r1_full_parity_v6.json 2> gpurun_out/r1_full_parity_v6.err; echo full-parity $?; tail -c 400 gpurun_out/r1_full_parity_v6.json
timeout 600 python tools/fuzz_gpu.py --cases 4000 --seed 11 > gpurun_out/fuzz_gpu_4000.json 2> gpurun_out/fuzz_gpu_4000.err; echo fuzz $?; tail -c 500 gpurun_out/fuzz_gpu_4000.json
