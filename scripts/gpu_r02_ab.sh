#!/bin/bash
# same box A/B of library variants (scripts/build_variants.sh): current vs no peer stores vs no probes vs neither
mkdir -p gpurun_out
for rep in 1 2; do
for v in "" nopeers noprobe plain; do
  if [ -n "$v" ]; then export DLWP_B200_LIB=$PWD/dlwp_b200/csrc/libdlwp_b200_$v.so; else unset DLWP_B200_LIB; fi
  echo "== variant '${v:-current}' rep $rep"
  timeout 120 python scripts/prof_tc.py --batch 256 --iters 20 2>&1 | tail -1
done
done > gpurun_out/r02_variants_ab.txt 2>&1
unset DLWP_B200_LIB
cat gpurun_out/r02_variants_ab.txt
timeout 900 python -m pytest tests/test_estimator_gpu.py tests/test_conv_gpu.py -q -x 2>&1 | tail -15
