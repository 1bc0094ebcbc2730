#!/bin/bash
# One `ncu --set full` capture per hot kernel of the default bench step (second step of the run), exported on the box as raw CSV +
# details text into gpurun_out/ (the .ncu-rep files with sources exceed what travels back).  Usage: bash profiles/ncu_capture.sh [tag]
tag=${1:-r02}
export CRM_BENCH_FIXED_WARMUP=1
capture() {   # kernel regex, launches to skip, output name
  ncu --set full --clock-control none -k "regex:$1" --launch-skip "$2" --launch-count 1 -o /tmp/ncu_$3 -f \
      python bench.py --steps 1 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > /dev/null 2> /tmp/ncu_$3.err
  ncu -i /tmp/ncu_$3.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_$3_raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$3.ncu-rep --page details > gpurun_out/${tag}_ncu_$3.txt 2>/dev/null
  ls -la gpurun_out/${tag}_ncu_$3.txt
}
capture '^oz_mma_kernel$' 2 oz_mma_kernel
capture 'oz_slice_kernel' 1 oz_slice_kernel
capture 'crm_score' 1 crm_score_kernel
capture 'crm_fit_kernel' 1 crm_fit_kernel
capture 'crm_sytrd_kernel' 1 crm_sytrd_kernel
capture 'oz_genotype_kernel' 1 oz_genotype_kernel
