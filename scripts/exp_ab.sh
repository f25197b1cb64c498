#!/bin/bash
# A/B of the force kernel variants selectable through NBB200_EXP (development aid; run under gpurun): parity tests first, then M1 and DHFR timings
o=gpurun_out; mkdir -p $o
python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -4 > $o/ab_tests.log
for e in ${EXPS:-0 8}; do
  echo "== NBB200_EXP=$e" >> $o/ab.log
  NBB200_EXP=$e python scripts/variant_bench.py m1 128x4 >> $o/ab.log 2>&1
  NBB200_EXP=$e python scripts/jac_probe.py >> $o/ab.log 2>&1
done
cat $o/ab_tests.log $o/ab.log
