#!/bin/bash
# ncu --set full of the fit kernel (with the table of bracket points) and of the genotype conversion kernel, second step of the bench command
tag=${1:-r02b}
export CRM_BENCH_FIXED_WARMUP=1
for spec in "crm_fit_kernel:1:crm_fit_kernel" "oz_genotype_kernel:1:oz_genotype_kernel"; do
  IFS=: read -r rx skip name <<< "$spec"
  ncu --set full --clock-control none -k "regex:$rx" --launch-skip "$skip" --launch-count 1 -o /tmp/ncu_$name -f \
      python bench.py --steps 1 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > /dev/null 2> /tmp/ncu_$name.err
  ncu -i /tmp/ncu_$name.ncu-rep --page details > gpurun_out/${tag}_ncu_$name.txt 2>/dev/null
  grep -E "^    (Duration|DRAM Throughput|Compute \(SM\) Throughput|Registers Per Thread|Achieved Occupancy|Executed Ipc Active)" gpurun_out/${tag}_ncu_$name.txt
done
