timeout 300 python tools/dbg_fused.py c3 > gpurun_out/r2p_dbg.log 2>&1
tail -6 gpurun_out/r2p_dbg.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -5 gpurun_out/r2q_pytest.log
GMD_FUSED=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2q_bench.err | grep "^{" > gpurun_out/r2q_bench_n1_fused1.json
cut -c1-300 gpurun_out/r2q_bench_n1_fused1.json
timeout 300 python bench.py --trace gpurun_out/r2q_trace_n1 > /dev/null 2>> gpurun_out/r2q_bench.err
tail -3 gpurun_out/r2q_bench.err
