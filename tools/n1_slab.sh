python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2k_pytest_gpu.txt; cat gpurun_out/r2k_pytest_gpu.txt
python bench.py --slabs --steps 10 --warmup 3 > gpurun_out/r2k_slab1.json 2> gpurun_out/r2k_slab1.err; tail -c 300 gpurun_out/r2k_slab1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2k_slab1.json') if l.startswith('{')][-1])
print(1, d['value']/1e9, d['ms_per_step'], d.get('phase_ms_rank0'), d['e2e']['value']/1e9)"
