(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r3b_pytest.log 2>&1; tail -6 gpurun_out/r3b_pytest.log
cp profiles/parity_r2.json gpurun_out/parity_r2.json 2>/dev/null
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r3b_bench_n1.json 2> gpurun_out/r3b_bench.err; cut -c1-200 gpurun_out/r3b_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3b_bench_ref.json 2>> gpurun_out/r3b_bench.err; cut -c1-300 gpurun_out/r3b_bench_ref.json
