#!/bin/bash
# The 8-GPU measurement pass (one `gpurun --gpus 8` call): weak scaling of the headline config at 1/2/4/8, BASELINE configs 3/4/5
# at 8 GPUs, the strong-scaling sweep of config 5 (semi_supervised, global batch 16384) at 1/2/4/8, and the 2-GPU plumbing check.
# Every line is one bench.py JSON line; outputs go to gpurun_out/<tag>_*.json.
tag=${1:-r2_scale8}
out=gpurun_out
S="--steps 20 --warmup 5 --no-cpu-baseline --no-module-path"
port=29600
run() {   # run <n_gpus> <name> <bench args...>
    local n=$1 name=$2; shift 2
    port=$((port + 1))
    if [ "$n" = "1" ]; then
        timeout 300 python bench.py --gpus 1 $S "$@" 2> $out/${tag}_${name}.err | tail -1 > $out/${tag}_${name}.json
    else
        timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n $S "$@" \
            2> $out/${tag}_${name}.err | tail -1 > $out/${tag}_${name}.json
    fi
    python - "$out/${tag}_${name}.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print(f"{sys.argv[2]:28s} n={d['n_gpus']} B/gpu={d['config']['per_gpu_batch']:6d} {d['ms_per_step']:8.3f} ms  {d['value']:12.0f} samples/s  e2e {d['e2e']['value']:12.0f}  {d['scaling']}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
# plumbing first (cheap): DP == single process bit for bit, replicas identical, DP CUDA-graph replay
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29599 tests/manual/dp_check.py > $out/${tag}_dp_check.log 2>&1
grep "mode=\|DP_CHECK\|replay" $out/${tag}_dp_check.log
# weak scaling, headline config (default mode, per-GPU batch 1024)
for n in 8 4 2 1; do run $n weak_default_n$n; done
# BASELINE config 3: mse mode on 8 GPUs; config 4: infonce with per-rank batch 8192 on 8 GPUs
run 8 mse_n8 --mode mse
run 8 infonce_b8192_n8 --mode infonce --batch 8192
# BASELINE config 5: semi_supervised, GLOBAL batch 16384, strong scaling 8/4/2/1
for n in 8 4 2 1; do run $n semi_g16384_n$n --mode semi_supervised --global-batch 16384; done
# small per-GPU batch through the CUDA graph (collectives captured) at 8 GPUs
run 8 graph_b128_n8 --batch 128 --graph
run 8 eager_b128_n8 --batch 128
