#!/bin/bash
# cooperative builder A/B (development aid; run under gpurun)
o=gpurun_out; mkdir -p $o; rm -f $o/coop.log
python -m pytest tests -m gpu -q 2>&1 | tail -30 > $o/coop_tests.log
for v in "NBB200_COOP=0" "NBB200_COOP=1"; do
  echo "== $v" >> $o/coop.log
  env $v python scripts/jac_probe.py >> $o/coop.log 2>&1
  env $v python scripts/md_probe.py >> $o/coop.log 2>&1
done
tail -n 6 $o/coop_tests.log; cat $o/coop.log
