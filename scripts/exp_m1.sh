#!/bin/bash
# GPU tests, M1 step timing (one GPU) and one rank's share at 1/8 (development aid; run under gpurun)
o=gpurun_out; mkdir -p $o
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/variant_bench.py m1 128x4
python scripts/part_probe.py 3/8
python scripts/jac_probe.py
