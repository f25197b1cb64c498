#!/bin/bash
# GPU test suite + the MD / JAC probes of the current build (development aid; run under gpurun)
o=gpurun_out; mkdir -p $o; rm -f $o/quick.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $o/quick_tests.log
for v in ${VARIANTS:-A=1}; do
  echo "== $v" >> $o/quick.log
  env $v python scripts/jac_probe.py >> $o/quick.log 2>&1
  env $v python scripts/md_probe.py >> $o/quick.log 2>&1
done
cat $o/quick_tests.log $o/quick.log
