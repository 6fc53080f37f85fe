#!/bin/bash
# compute-sanitizer over the kernels changed this round (small shapes only): memcheck + racecheck
O=gpurun_out
K='add_golden or near_tie or lean_add_routes or tree_equals_scan or retessellate or zero_copy or pipelined'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_archives.py tests/test_gpu_round2.py -m gpu -q -x -k "$K" > $O/r02q_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" $O/r02q_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "near_tie or lean_add_routes" > $O/r02q_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/r02q_racecheck.log | tail -8
