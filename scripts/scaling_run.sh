#!/bin/bash
# bench.py at N = 2, 4, 8 on one box (run under gpurun --gpus 8); N = 1 comes from collect_evidence.sh
o=gpurun_out; tag=${1:-r02}
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err
done
python bench.py --steps 10 --warmup 3 --no-cpu --no-jac > $o/${tag}_bench_n1_samebox.json 2>/dev/null
