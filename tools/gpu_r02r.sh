#!/bin/bash
# round 2, call R: compute-sanitizer on the kernels written this round (pair-lane classify, tile-major hysteresis,
# dilation by lane shuffles, hash rows through shared memory, persistent fused pass)
O=gpurun_out/r02r; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
T1="tests/test_gpu_parity.py::test_edge_intermediates_match_cv2 tests/test_gpu_parity.py::test_frame_hashes_match_cv2 tests/test_gpu_parity.py::test_integer_sums_any_shape tests/test_gpu_parity.py::test_halo_shards_equal_serial"
timeout 1200 $CS --tool memcheck --error-exitcode 9 python -m pytest $T1 -q -m gpu -x > $O/memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/memcheck.txt; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck.txt | tail -4
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_parity.py::test_batched_scene_manager_matches_reference_golden" -q -m gpu -x -k "edges or hash or hist" > $O/memcheck_goldens.txt 2>&1; echo "memcheck goldens rc=$?" | tee -a $O/memcheck_goldens.txt; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_goldens.txt | tail -4
timeout 900 $CS --tool racecheck python -m pytest "tests/test_gpu_parity.py::test_frame_hashes_match_cv2" "tests/test_gpu_parity.py::test_edge_intermediates_match_cv2" -q -m gpu -x > $O/racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a $O/racecheck.txt; grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck.txt | tail -4
timeout 900 $CS --tool initcheck python -m pytest "tests/test_gpu_parity.py::test_edge_intermediates_match_cv2" -q -m gpu -x > $O/initcheck.txt 2>&1; echo "initcheck rc=$?" | tee -a $O/initcheck.txt; grep -E "ERROR SUMMARY|passed|failed" $O/initcheck.txt | tail -4
ls -la $O
