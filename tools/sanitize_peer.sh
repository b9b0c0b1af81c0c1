#!/bin/bash
# compute-sanitizer over a 2-rank peer-memory run of one small wide-halo case (tests/multi_gpu_check.py case 0: csp2,
# upwind, filter rows, 120x61, one model step): the release / acquire halo protocol, the fused peer stores of the S3a
# launches and the one-shot all-reduce all run.  Twice: through the three-sweep path (what bands this small pick) and,
# GMD_FUSED=1, through the fused predict_correct kernel k_pc with its peer stores (synccheck not there: see pc_bar in
# gmd_pc.cuh and tools/sanitize_fused.sh).  Usage (2 GPUs):  bash tools/sanitize_peer.sh > profiles/..._sanitizer_peer.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
export GMD_CASES=0 GMD_KW='{"nsteps": 1}'
for spec in "0 memcheck" "0 racecheck" "0 synccheck" "1 memcheck" "1 racecheck"; do
  set -- $spec
  export GMD_FUSED=$1
  echo "== GMD_FUSED=$1 compute-sanitizer --tool $2 --target-processes all  $TR tests/multi_gpu_check.py peer"
  timeout 900 compute-sanitizer --tool $2 --target-processes all --error-exitcode 9 $TR tests/multi_gpu_check.py peer 2>&1 |
    grep -E "ranks, |ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error|error" | grep -v "OMP_NUM" | sort | uniq -c | head -20
  echo "exit code ${PIPESTATUS[0]}"
done
