# usage (under gpurun): bash scripts/gpu_quick.sh <tag>   -> GPU tests + both headline workloads, kernel times
mkdir -p gpurun_out
TAG=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_${TAG}.log
for WL in ${WLS:-c2_plate_qeph_1m c5_brick_slab_2m}; do
  python bench.py --workload $WL --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_${TAG}_${WL}.err | tail -1 > gpurun_out/bench_${TAG}_${WL}.json
  python -c "import sys,json; d=json.load(open('gpurun_out/bench_${TAG}_${WL}.json')); print('$WL', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'node', d['roofline']['node_kernel'], 'e2e %.4g'%d['e2e']['value'], d['clocks'])"
done
