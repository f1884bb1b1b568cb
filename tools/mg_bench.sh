#!/bin/bash
# usage: tools/mg_bench.sh "8 4 2"   -- multi-GPU parity tests + config 5 strong-scaling runs (needs that many GPUs)
python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
for n in ${1:-8 4 2}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2j_slab$n.json 2> gpurun_out/r2j_slab$n.err
  tail -c 300 gpurun_out/r2j_slab$n.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2j_slab$n.json') if l.startswith('{')][-1])
print($n, d['value']/1e9, d['ms_per_step'], d['overlap']['ms_per_step_without_overlap'], d['phase_ms_rank0'], d['e2e']['value']/1e9)"
done
