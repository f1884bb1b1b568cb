#!/bin/bash
# one GPU: the default bench invocation (what the driver runs), wall clock
s=$(date +%s)
python bench.py > gpurun_out/r2k_bench_default.json 2> gpurun_out/r2k_bench_default.err
echo "wall $(( $(date +%s) - s )) s"; tail -c 200 gpurun_out/r2k_bench_default.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2k_bench_default.json') if l.startswith('{')][-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['cpu_baseline'], list(d['configs'].keys()), d['configs']['5_one_gpu'].get('ms_per_step'))"
