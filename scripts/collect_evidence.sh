#!/bin/bash
# Round evidence on one B200 (run under gpurun): test log, bench line, launch lists and ncu --set full captures of the kernels the
# round's numbers come from.  Everything lands in gpurun_out/<tag>_*; the summaries are made afterwards with scripts/summarize_ncu.py.
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $o/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1
python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err
python bench.py --impl reference --steps 1 --warmup 0 > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $o/${tag}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-jac > $o/${tag}_launches_bench.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $o/${tag}_launches_m1_norebuild.csv python scripts/profile_run.py m1 3 norebuild > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $o/${tag}_launches_dhfr.csv python scripts/profile_run.py dhfr 4 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $o/${tag}_launches_md_dhfr.csv python scripts/profile_run.py dhfr_mm 20 md > /dev/null 2>&1
FULL="$NCU --set full --import-source on -f"
$FULL -k regex:k_cluster_forces -s 2 -c 1 -o $o/${tag}_forces_m1 python scripts/profile_run.py m1 3 > $o/${tag}_prof.log 2>&1
$FULL -k regex:k_build_tiles -s 2 -c 1 -o $o/${tag}_build_m1 python scripts/profile_run.py m1 3 >> $o/${tag}_prof.log 2>&1
$FULL -k "regex:k_cluster_forces|k_prune" -s 2 -c 2 -o $o/${tag}_norebuild_m1 python scripts/profile_run.py m1 3 norebuild >> $o/${tag}_prof.log 2>&1
$FULL -k "regex:k_cluster_forces|k_build_tiles" -s 6 -c 2 -o $o/${tag}_dhfr python scripts/profile_run.py dhfr 4 >> $o/${tag}_prof.log 2>&1
