#!/bin/bash
# K0 capture + launch list only (the rest of profiles/ncu_capture.sh was taken earlier in the round)
tag=${1:-r02b}
export CRM_BENCH_FIXED_WARMUP=1
ncu --set full --clock-control none -k "regex:^oz_mma_kernel$" --launch-skip 1 --launch-count 1 -o /tmp/ncu_k0 -f \
    python bench.py --steps 1 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > /dev/null 2> /tmp/ncu_k0.err
ncu -i /tmp/ncu_k0.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_oz_mma_kernel_raw.csv 2>/dev/null
ncu -i /tmp/ncu_k0.ncu-rep --page details > gpurun_out/${tag}_ncu_oz_mma_kernel.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > /dev/null 2> /tmp/ncu_launches.err
ls -la gpurun_out/
