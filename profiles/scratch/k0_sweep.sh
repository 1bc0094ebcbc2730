#!/bin/bash
# A/B of the int8 contraction variants on the compact basis (CRM_TRACE phases of profiles/step_trace.py)
for cfg in "" "CRM_OZ_NGROUP=4" "CRM_OZ_NGROUP=16" "CRM_OZ_NGROUP=40" "CRM_INT8_MMA=2cta"; do
  echo "== ${cfg:-default}"
  env $cfg CRM_TRACE=1 python profiles/step_trace.py --reps 4 2>&1 | grep "int8 rotation" | tail -2 | sed 's/.*digit planes/digit planes/'
done
