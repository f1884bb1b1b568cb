#!/bin/bash
# 2 GPUs: compute-sanitizer memcheck over the 2-GPU parity test (peer-memory link, strided append,
# gather / sweep-only layer calls, both overlap modes)
timeout 420 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 \
    python -m pytest tests/test_multigpu.py -m gpu -x -q -k two_gpu > gpurun_out/r2k_mg_memcheck.log 2>&1
echo "rc=$?"; grep -c "ERROR SUMMARY" gpurun_out/r2k_mg_memcheck.log; grep "ERROR SUMMARY\|passed\|failed\|Error" gpurun_out/r2k_mg_memcheck.log | sort | uniq -c | head -20
