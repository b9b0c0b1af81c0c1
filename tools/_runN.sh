N=$1; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for pbr in "$@"; do
timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --polar-band-rows $pbr --no-decomposition-check 2>> gpurun_out/r2z_n$N.err | grep "^{" > gpurun_out/r2z_bench_n${N}_pbr$pbr.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2z_bench_n${N}_pbr$pbr.json"))
print($pbr, d["ms_per_step"], d["parallel"], d["conservation"])
PY
done
grep -v "OMP_NUM_THREADS\|\*\*\*\*\*\|^$\|NCCL version" gpurun_out/r2z_n$N.err | tail -3
