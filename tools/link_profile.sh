#!/bin/bash
# one GPU: the exchange kernels of csrc/link.cu (in-process links): time, ncu capture, sanitizer
python tools/link_times.py 20 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:k_link_classify_send -s 2 -c 1 -o gpurun_out/r2_link_send python tools/link_times.py 4 > gpurun_out/r2_link_ncu.log 2>&1; tail -2 gpurun_out/r2_link_ncu.log
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_link.py -m gpu -q -x > gpurun_out/r2_link_racecheck.log 2>&1; tail -3 gpurun_out/r2_link_racecheck.log
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_link.py -m gpu -q -x > gpurun_out/r2_link_memcheck.log 2>&1; tail -3 gpurun_out/r2_link_memcheck.log
