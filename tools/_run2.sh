TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --workload sg_band8 --steps 20 --warmup 5 --no-decomposition-check > gpurun_out/r2f_band8_cap.json 2> gpurun_out/r2f.err
cut -c1-330 gpurun_out/r2f_band8_cap.json
GMD_NO_CAP=1 $TR bench.py --gpus 2 --workload sg_band8 --steps 20 --warmup 5 --no-decomposition-check > gpurun_out/r2f_band8_nocap.json 2>> gpurun_out/r2f.err
cut -c1-330 gpurun_out/r2f_band8_nocap.json
$TR bench.py --gpus 2 --workload sg_band8 --trace gpurun_out/r2f_trace_band8_cap 2>> gpurun_out/r2f.err
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_n2_cap.json 2>> gpurun_out/r2f.err
cut -c1-330 gpurun_out/r2f_n2_cap.json
python -m pytest tests -m gpu -x -q -k "band_decomposition" > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
grep -v "OMP_NUM_THREADS\|\*\*\*\*\*\|^$" gpurun_out/r2f.err | tail -5
