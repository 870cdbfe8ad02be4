#!/bin/bash
# Runs the parity tests with every alternative code path forced through its environment knob:
# table kernel only, no multi-layout launches, in-place launches always / never lazy, no
# auto-crop, register-staged instead of TMA-staged overlay, interleaved chunk order, and the two
# other host-frame modes. Every line must end in "passed" (1296 tests each on a B200).
#   gpurun -- bash tools/knob_matrix.sh
K="rectangles_match_oracle or golden or fuzz or alpha_sweep or scaled_rectangles or regions or unaligned"
for v in FLUC_TTMLBLEND_GROUPS=0 FLUC_TTMLBLEND_MULTI=0 FLUC_TTMLBLEND_LAZY=1 FLUC_TTMLBLEND_LAZY=0 \
         FLUC_TTMLBLEND_AUTOCROP=0 FLUC_TTMLBLEND_BULK=0 FLUC_TTMLBLEND_LANES=7 \
         FLUC_TTMLBLEND_HOST_MODE=0 FLUC_TTMLBLEND_HOST_MODE=2; do
  echo "== $v"
  env $v timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_regions.py \
      -q -x -k "$K" 2>&1 | tail -1
done
