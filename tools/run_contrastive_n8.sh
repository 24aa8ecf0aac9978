#!/bin/bash
# 8-GPU lines of the stand-alone contrastive steps (BASELINE config 4 read literally: large-batch InfoNCE / SimCLR on 8 x B200) + the DP check
out=gpurun_out; tag=${1:-r2zb}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29821 tests/manual/dp_check_contrastive.py > $out/${tag}_dp_check_contrastive.log 2>&1
grep -E "kind=|DP_CHECK" $out/${tag}_dp_check_contrastive.log
port=29830
for cfg in "contrastive_infonce 8192" "contrastive_simclr 2048" "contrastive_infonce 1024"; do
  set -- $cfg; port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --kind $1 --batch $2 --steps 10 --warmup 3 2> $out/${tag}_$1_b$2_n8.err | tail -1 > $out/${tag}_$1_b$2_n8.json
  python -c "
import json,sys
d=json.load(open('$out/${tag}_$1_b$2_n8.json')); print('$1', 'B/gpu', d['config']['per_gpu_batch'], 'n', d['n_gpus'], round(d['ms_per_step'],3), 'ms', round(d['value']), 'samples/s', 'e2e', round(d['e2e']['value']))"
done
timeout 300 python bench.py --kind contrastive_simclr --batch 2048 --steps 10 --warmup 3 2>> $out/${tag}.err | tail -1 > $out/${tag}_contrastive_simclr_b2048_n1.json
python -c "
import json
d=json.load(open('$out/${tag}_contrastive_simclr_b2048_n1.json')); print('simclr n1 B 2048', round(d['ms_per_step'],3), round(d['value']))"
