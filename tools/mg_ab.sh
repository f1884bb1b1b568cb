#!/bin/bash
# usage: tools/mg_ab.sh N   -- A/B of the overlapped multi-GPU step: side-stream priority x pipelining
n=${1:-2}
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  PNB_SLAB_PRIORITY=$1 PNB_SLAB_PIPELINE=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2k_ab_${n}_$1$2.json 2> gpurun_out/r2k_ab.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2k_ab_${n}_$1$2.json') if l.startswith('{')][-1])
print('prio $1 pipeline $2:', round(d['ms_per_step'],3), 'no-ovl', round(d['overlap']['ms_per_step_without_overlap'],3), 'split', round(d['overlap']['ms_per_step_mode_split'],3), 'e2e ms', round(d['e2e']['ms_per_step'],2), d['step_ms_rank0'])"
done
