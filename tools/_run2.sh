TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for b in 1 2 3 4; do
echo ROWS_B $b
GMD_PDL=1 GMD_ROWS_PER_CTA_B=$b $TR bench.py --gpus 2 --workload sg_band8 --steps 30 --warmup 5 --no-decomposition-check 2>> gpurun_out/r2j.err | cut -c1-330
done
grep -v "OMP_NUM_THREADS\|\*\*\*\*\*\|^$" gpurun_out/r2j.err | tail -5
