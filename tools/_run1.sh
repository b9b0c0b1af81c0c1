timeout 600 compute-sanitizer --tool synccheck --print-limit 3 python __graft_entry__.py smoke 2>&1 | grep -E "smoke ok|ERROR SUMMARY|Barrier error" | sort | uniq -c | head
GMD_FUSED=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2x_bench.err | grep "^{" | cut -c1-250
