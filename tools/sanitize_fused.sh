#!/bin/bash
# compute-sanitizer over a small one-GPU case that runs the fused predict_correct kernel k_pc (72x37 mountain flow, csp2:
# slow / fast passes, all three deferred-update variants, polar-side chain beside it): racecheck for its shared-memory
# rings, memcheck, and synccheck on the build with the unaligned tick barrier (see pc_bar in gmd_pc.cuh).
# Usage (1 GPU):  bash tools/sanitize_fused.sh > profiles/..._sanitizer_fused.txt
set -u
cd "$(dirname "$0")/.."
mkdir -p build_exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DGMD_PC_UNALIGNED_BAR=1 -shared \
  -o build_exp/libgmd_unaligned_bar.so gamil_dycore_b200/csrc/gmd.cu -ldl || exit 1
cat > /tmp/_san_case.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
os.environ['GMD_FUSED'] = '1'   # a grid this small would not pick the fused kernel by itself
import gamil_dycore_b200 as gmd
kw = dict(num_lon=72, num_lat=37, time_step_size=600.0, subcycles=4, split_scheme="csp2", zonal_tend_filter_cutoff_wavenumber=[4, 4, 4])
u, v, gd, ghs = gmd.initial_condition("mountain_zonal_flow", 72, 37)
d = gmd.Dycore(gmd.Config(**kw), kind=sys.argv[1])
assert d.fused_rows()[1] > d.fused_rows()[0]
d.set_state(u, v, gd, ghs); d.run_init(); d.step(2)
m, e, _ = d.diag()
print(f"case ok: fused rows {d.fused_rows()}, {d.kernel_launches()} launches, mass {m:.15e} energy {e:.15e}")
PY
for spec in "racecheck fast" "memcheck fast" "synccheck $PWD/build_exp/libgmd_unaligned_bar.so"; do
  set -- $spec
  echo "== compute-sanitizer --tool $1  (libgmd: $2)"
  timeout 900 compute-sanitizer --tool $1 --error-exitcode 9 python /tmp/_san_case.py $2 2>&1 |
    grep -E "case ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Barrier error|rror" | sort | uniq -c | head -8
  echo "exit code ${PIPESTATUS[0]}"
done
