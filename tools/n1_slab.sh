#!/bin/bash
# one GPU: smoke, the N = 1 line of the slab code path (config 5 on one GPU), the default bench line
python tools/smoke_only.py 2>&1 | tail -2
python bench.py --slabs --steps 10 --warmup 3 > gpurun_out/r2k_slab1.json 2> gpurun_out/r2k_slab1.err; tail -c 300 gpurun_out/r2k_slab1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2k_slab1.json') if l.startswith('{')][-1])
print(1, d['value']/1e9, d['ms_per_step'], d.get('phase_ms_rank0'), d['e2e']['value']/1e9)"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; tail -c 300 gpurun_out/r2k_bench_n1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2k_bench_n1.json') if l.startswith('{')][-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['roofline']['frac'], d['clocks'])"
