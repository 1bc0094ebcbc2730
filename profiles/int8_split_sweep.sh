#!/bin/bash
# rasterisation sweep of the fused int8 contraction (SNP tiles per group; CTA pairs vs single CTAs) at bench size
for cfg in "2cta 2" "2cta 8" "2cta 16" "2cta 40" "1cta 8" "1cta 16" "1cta 2"; do
  set -- $cfg
  echo "== $1 ngroup=$2"
  CRM_INT8_MMA=$1 CRM_OZ_NGROUP=$2 timeout 120 python profiles/int8_split_bench.py 100000 21504 10000 fused-only 2>&1 | tail -1
done
