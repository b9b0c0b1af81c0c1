GMD_FUSED=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2u_bench.err | grep "^{" > gpurun_out/r2u_bench_n1_fused1.json
cut -c1-300 gpurun_out/r2u_bench_n1_fused1.json
timeout 300 python tools/dbg_fused.py c3 c1 > gpurun_out/r2u_dbg.log 2>&1
grep "after 5" gpurun_out/r2u_dbg.log | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pc -s 14 -c 1 -o gpurun_out/r2u_pc python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2u_ncu.log 2>&1
