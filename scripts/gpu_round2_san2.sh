#!/bin/bash
# compute-sanitizer on the block-cooperative long-range tiers (k_geodesic_cta: shared-memory workspace on the default-executable
# shape, global-memory workspace behind css_distance on the whole mesh)
mkdir -p gpurun_out
SEL='default_executable or (self_consistency_on_the_gpu and cfg1)'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_real_meshes.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_cta_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_cta_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_real_meshes.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_cta_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_cta_synccheck.log | tail -3
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_real_meshes.py -m gpu -q -x -k "default_executable" > gpurun_out/r2_sanitizer_cta_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_cta_racecheck.log | tail -3
grep -E "Race reported|hazard" gpurun_out/r2_sanitizer_cta_racecheck.log | sed -E 's/0x[0-9a-f]+//g' | sort | uniq -c | sort -rn | head -30
