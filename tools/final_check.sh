#!/bin/bash
# round-2 validation on one B200: full GPU test suite, bench line, launch list, ncu captures of
# the two hot kernels, compute-sanitizer on the new kernels.  Outputs under gpurun_out/.
o=gpurun_out
python tools/smoke_only.py 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $o/r2_pytest_gpu.txt; cat $o/r2_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > $o/r2_bench_n1.json 2> $o/r2_bench_n1.err; tail -c 300 $o/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-configs > $o/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_sweep_flat$ -s 2 -c 1 -o $o/r2_wcsph_flat_final \
    python tools/kernel_times.py --closures wcsph --reps 1 > $o/r2_ncu_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_bucket_scatter$ -s 3 -c 1 -o $o/r2_update_final \
    python tools/kernel_times.py --closures wcsph --reps 2 > $o/r2_ncu_update.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "wcsph_parity or nbody_parity or stream_ordered or live_neighbor or two_sets_tile or periodic_lists or float64_search or prefilter_never or edge_cases" \
    > $o/r2_memcheck.log 2>&1; tail -4 $o/r2_memcheck.log
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "(wcsph_parity and fast and 24) or (two_sets_tile and False) or (periodic_lists and 3-24)" \
    > $o/r2_racecheck.log 2>&1; tail -4 $o/r2_racecheck.log
