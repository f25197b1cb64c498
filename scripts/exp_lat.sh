#!/bin/bash
# latency experiments at JAC size (development aid; run under gpurun): GPU tests, then the DHFR call timings and the MD loop with and without
# the variants selectable through the environment
o=gpurun_out; mkdir -p $o; rm -f $o/lat.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $o/lat_tests.log
for v in "" "NBB200_NO_PUBLISH=1" "NBB200_NO_FUSE=1" "NBB200_NO_PUBLISH=1 NBB200_NO_FUSE=1"; do
  echo "== $v" >> $o/lat.log
  env $v python scripts/jac_probe.py >> $o/lat.log 2>&1
  env $v python scripts/md_probe.py >> $o/lat.log 2>&1
done
cat $o/lat_tests.log $o/lat.log
