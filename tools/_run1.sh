export GMD_PARITY_OUT=gpurun_out/parity.json
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r3f_pytest.log 2>&1; tail -6 gpurun_out/r3f_pytest.log
unset GMD_PARITY_OUT
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r3f_bench_n1.json 2> gpurun_out/r3f_bench.err; cut -c1-200 gpurun_out/r3f_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3f_bench_ref.json 2>> gpurun_out/r3f_bench.err; cut -c1-200 gpurun_out/r3f_bench_ref.json
for w in jz_0.25deg rh_0.05deg; do timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r3f_bench_n1_$w.json 2>> gpurun_out/r3f_bench.err; cut -c1-200 gpurun_out/r3f_bench_n1_$w.json; done
timeout 300 python bench.py --trace gpurun_out/r3f_trace_n1 > /dev/null 2>> gpurun_out/r3f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_polar_lean -s 20 -c 3 -o gpurun_out/r3f_polar python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r3f_ncu.log 2>&1
tail -2 gpurun_out/r3f_ncu.log
