N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus $N --steps 40 --warmup 5 2>> gpurun_out/r3h_n$N.err | grep "^{" > gpurun_out/r3h_bench_n${N}.json
python - <<PY
import json
d=json.load(open("gpurun_out/r3h_bench_n${N}.json"))
print(d["ms_per_step"], d["value"]/1e9, d["parallel"], d.get("decomposition_check",{}).get("decomposition_max_rel"))
PY
