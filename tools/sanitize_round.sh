#!/bin/bash
# compute-sanitizer evidence (run on the GPU box through gpurun): memcheck over the golden GPU parity tests, then
# racecheck and synccheck over one EGNO and one SEGNO golden case.  Logs go to gpurun_out/ (copy digests to profiles/).
cd ${GRAFT_REPO_ROOT:-.}
S=/usr/local/cuda/bin/compute-sanitizer
K='test_egno_matches_reference_golden or test_segno_matches_reference_golden or test_edge_tile_building_block_vs_oracle or test_blocked_selector_walk'
timeout 900 $S --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_memcheck.log
K2='test_egno_matches_reference_golden and egno_n5_t8 or test_segno_matches_reference_golden and segno_n5_t10'
timeout 600 $S --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K2" > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_racecheck.log
timeout 600 $S --tool synccheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K2" > gpurun_out/r02_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/r02_synccheck.log
tail -4 gpurun_out/r02_memcheck.log gpurun_out/r02_racecheck.log gpurun_out/r02_synccheck.log
