#!/bin/bash
# Runs on an N-GPU box (gpurun --gpus N): both arms of bench.py under torchrun, outputs into gpurun_out/.
#   tools/bench_multi.sh N [steps]
set -u
N=$1; K=${2:-200}; O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps $K --warmup 5 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err
echo "bench n=$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/r02_bench_n${N}_reference.json 2> $O/r02_bench_n${N}_reference.err
echo "reference n=$N rc=$?"
tail -c 600 $O/r02_bench_n$N.err
head -c 1500 $O/r02_bench_n$N.json
