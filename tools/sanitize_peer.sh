#!/bin/bash
# compute-sanitizer over a 2-rank peer-memory run of one small wide-halo case (tests/multi_gpu_check.py case 0: csp2,
# upwind, filter rows, 120x61, one model step): the release / acquire halo protocol, the fused peer stores of the S3a
# launches and the one-shot all-reduce all run.  Usage (2 GPUs):  bash tools/sanitize_peer.sh > profiles/..._sanitizer.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
export GMD_CASES=0 GMD_KW='{"nsteps": 1}'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool --target-processes all  $TR tests/multi_gpu_check.py peer"
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 $TR tests/multi_gpu_check.py peer 2>&1 |
    grep -E "ranks, |ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error|error" | grep -v "OMP_NUM" | sort | uniq -c | head -20
  echo "exit code ${PIPESTATUS[0]}"
done
