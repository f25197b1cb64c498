#!/bin/bash
# MD loop A/B (development aid; run under gpurun)
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "md_loop or langevin or verlet or bonded" 2>&1 | tail -3
for v in "A=1" "NBB200_MD_NO_PREDICTION=1"; do echo "== $v"; env $v python scripts/md_probe.py 2>&1; done
