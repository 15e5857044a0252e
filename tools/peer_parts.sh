#!/bin/bash
# Runs on an N-GPU box: tools/peer_tick_bench.py under the given environment strings, one line each.
#   tools/peer_parts.sh N "RTPP_TICK_PARTS=2" "RTPP_TICK_STATIC_PCT=50" ...
N=$1; shift
port=29600
for envs in "" "$@"; do
    port=$((port + 1))
    echo "== $envs"
    env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tools/peer_tick_bench.py 100 2>/dev/null | grep n_gpus | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fused_ms', round(d['fused_ms'],4), 'per-rank', [round(x,4) for x in d['fused_ms_per_rank']], 'match', d['all_ranks_match_single_gpu'])"
done
