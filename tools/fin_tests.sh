#!/bin/bash
# one GPU: the full GPU suite, smoke, and the default bench invocation (what the driver runs)
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2k_pytest_gpu.txt; cat gpurun_out/r2k_pytest_gpu.txt
python tools/smoke_only.py 2>&1 | tail -1
/usr/bin/time -v python bench.py > gpurun_out/r2k_bench_default.json 2> gpurun_out/r2k_bench_default.err; grep "Elapsed (wall" gpurun_out/r2k_bench_default.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2k_bench_default.json') if l.startswith('{')][-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['cpu_baseline'], list(d['configs'].keys()), d['configs']['5_one_gpu'].get('ms_per_step'))"
