#!/bin/bash
# usage: tools/mg_bench.sh "8 4 2" [notest]  -- multi-GPU parity tests + config 5 strong-scaling runs (needs that many GPUs)
if [ "$2" != "notest" ]; then python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3; fi
for n in ${1:-8 4 2}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2k_slab$n.json 2> gpurun_out/r2k_slab$n.err
  tail -c 300 gpurun_out/r2k_slab$n.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2k_slab$n.json') if l.startswith('{')][-1])
print($n, d['value']/1e9, d['ms_per_step'], {k:v for k,v in d['overlap'].items() if k!='what'}, d['e2e']['value']/1e9)
print('step by rank', d['step_ms_median_by_rank'])
for k,v in d['phase_ms_by_rank'].items(): print('  %-24s'%k, v)
for k,v in d['timed_loop_by_rank'].items(): print('  %-24s'%k, v)"
done
