# usage (under gpurun --gpus N): bash scripts/gpu_multi.sh N [workload]
mkdir -p gpurun_out
N=$1; WL=${2:-plate_small}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_domains_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --workload $WL 2>&1 | tail -3 | tee gpurun_out/bench_multi_${N}_${WL}.log
