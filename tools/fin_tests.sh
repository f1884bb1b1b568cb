python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2k_pytest_gpu.txt; cat gpurun_out/r2k_pytest_gpu.txt
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_link.py -m gpu -q -x > gpurun_out/r2_link_racecheck.log 2>&1; tail -3 gpurun_out/r2_link_racecheck.log
