#!/bin/bash
# Trimmed 8-GPU pass at the end of round 2 (current HEAD): weak scaling of the headline config at 8/4/2/1, the simple encoder family and
# BASELINE configs 3 / 5 at 8 GPUs, and the data-parallel plumbing check.  One bench.py JSON line each -> gpurun_out/<tag>_*.json
tag=${1:-r2y}
out=gpurun_out
S="--steps 20 --warmup 5 --no-cpu-baseline --no-module-path"
port=29700
run() {
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
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29699 tests/manual/dp_check.py > $out/${tag}_dp_check.log 2>&1
grep "mode=\|DP_CHECK\|replay" $out/${tag}_dp_check.log | head -12
for n in 8 4 2 1; do run $n weak_default_n$n; done
run 8 multi_simple_n8 --kind multi_simple
run 8 multi_cross_attention_n8 --kind multi_cross_attention --batch 512
run 8 mse_n8 --mode mse
run 8 semi_g16384_n8 --mode semi_supervised --global-batch 16384
cat $out/${tag}_*.json > $out/${tag}_scale8_lines.jsonl
