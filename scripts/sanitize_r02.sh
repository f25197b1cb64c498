#!/bin/bash
# compute-sanitizer over the kernels added in the second half of round 2 (run under gpurun); summaries into gpurun_out/r02_sanitizer.txt
o=gpurun_out/r02_sanitizer.txt; : > $o
run() { echo "## $*" >> $o; timeout 400 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" | tail -4 >> $o; }
run compute-sanitizer --tool racecheck python scripts/profile_run.py bala 2
run compute-sanitizer --tool memcheck python scripts/profile_run.py dhfr 2
run compute-sanitizer --tool racecheck python scripts/profile_run.py dhfr_mm 4 md
run compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "three_partitions or optimistic or projected_on_the_translation or native_md_loop_equals"
cat $o
