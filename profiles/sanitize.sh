#!/bin/bash
# compute-sanitizer memcheck + racecheck on the smoke run (small problem through every kernel of the default interaction path: host feeder,
# int8 transposition, digit planes, K0 incl. the split-K variant, warp score kernel, affine route).  Usage: bash profiles/sanitize.sh [tag]
tag=${1:-r02}
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_${tool}.txt 2>&1
  tail -4 gpurun_out/${tag}_sanitizer_${tool}.txt
done
