#!/bin/bash
# the GPU parity suite on the alternative paths behind the developer switches
mkdir -p gpurun_out
: > gpurun_out/r2_pytest_alt_paths.log
for sw in "CSS_STENCIL=0" "CSS_NO_GRAPH=1" "CSS_PDL=0" "CSS_WIN_HALF=0"; do
  echo "== $sw" >> gpurun_out/r2_pytest_alt_paths.log
  env $sw timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 >> gpurun_out/r2_pytest_alt_paths.log
done
cat gpurun_out/r2_pytest_alt_paths.log
