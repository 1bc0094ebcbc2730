#!/bin/bash
# One `ncu --set full` capture per hot kernel of the default bench step, exported on the box as raw CSV + details text into gpurun_out/
# (the .ncu-rep files with sources exceed what travels back), then the launch list of one step.  Usage: bash profiles/ncu_capture.sh [tag]
tag=${1:-r02b}
export CRM_BENCH_FIXED_WARMUP=1
capture() {   # kernel regex, launches to skip, output name, [extra environment]
  env $4 ncu --set full --clock-control none -k "regex:$1" --launch-skip "$2" --launch-count 1 -o /tmp/ncu_$3 -f \
      python bench.py --steps 1 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > /dev/null 2> /tmp/ncu_$3.err
  ncu -i /tmp/ncu_$3.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_$3_raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$3.ncu-rep --page details > gpurun_out/${tag}_ncu_$3.txt 2>/dev/null
  ls -la gpurun_out/${tag}_ncu_$3.txt
}
capture '^oz_mma_kernel$' 1 oz_mma_kernel                # one launch per step (the g2 Grams take oz_mma_splitk_kernel)
capture 'oz_slice_kernel' 7 oz_slice_kernel          # 4 launches per gene with a structured background: the 4th is the hK.E_l.E_j section
capture 'kr_expand_kernel' 1 kr_expand_kernel
capture 'crm_sytrd_kernel' 1 crm_sytrd_kernel
# fp64 route: DRAM bytes and duration of every crm_gemm_kernel launch (the rotation is the long one)
CRM_ROTATION=dmma ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:crm_gemm_kernel -c 80 --csv \
    --log-file gpurun_out/${tag}_fp64_route_gemm_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > /dev/null 2> /tmp/ncu_fp64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > /dev/null 2> /tmp/ncu_launches.err
ls -la gpurun_out/${tag}_launches_bench.csv
