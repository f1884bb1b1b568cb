#!/bin/bash
# 2 GPUs: compute-sanitizer memcheck around EACH rank of the 2-GPU parity test (peer-memory link,
# strided append, gather / sweep-only layer calls, both overlap modes)
o=gpurun_out
for r in 0 1; do
  RANK=$r WORLD_SIZE=2 MASTER_PORT=29541 timeout 500 compute-sanitizer --tool memcheck --print-limit 20 \
      python tools/mg_worker.py > $o/r2k_mg_memcheck_rank$r.log 2>&1 &
done
wait
for r in 0 1; do grep "ERROR SUMMARY\|worker finished\|Error\|Invalid" $o/r2k_mg_memcheck_rank$r.log | sort | uniq -c | head; done
