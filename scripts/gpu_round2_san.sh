#!/bin/bash
# round 2: compute-sanitizer on small parity cases (all stage-1 paths, walker vertex / border events, stride guard)
mkdir -p gpurun_out
SEL='(test_neighbours_distances_tangents_forces and (icosphere16 or torus60x24)) or test_open_mesh_boundary_rules or test_edge_cases or test_stride_guard or test_gpu_vertex_crossings or test_gpu_boundary_vertex'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_vertex_crossings.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_memcheck.log | tail -3
CSS_STENCIL=0 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_neighbours_distances_tangents_forces and (icosphere16 or torus60x24)" > gpurun_out/r2_sanitizer_memcheck_floodfill.log 2>&1
echo "memcheck floodfill rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_memcheck_floodfill.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_neighbours_distances_tangents_forces and torus60x24" > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_racecheck.log | tail -3
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_neighbours_distances_tangents_forces and torus60x24" > gpurun_out/r2_sanitizer_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_synccheck.log | tail -3
