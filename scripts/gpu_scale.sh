# usage (under gpurun --gpus N): bash scripts/gpu_scale.sh N workload [steps]
mkdir -p gpurun_out
N=$1; WL=$2; ST=${3:-100}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_${N}_${WL}.csv &
SMI=$!
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps $ST --warmup 10 --workload $WL 2>gpurun_out/scale_${N}_${WL}.err | tail -1 | tee gpurun_out/scale_${N}_${WL}.json
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $ST --warmup 10 --workload $WL 2>gpurun_out/scale_${N}_${WL}.err | tail -1 | tee gpurun_out/scale_${N}_${WL}.json
fi
kill $SMI
tail -3 gpurun_out/scale_${N}_${WL}.err
