# usage (under gpurun): bash scripts/gpu_sweep.sh <workload> lib1 lib2 ...   -> kernel times per variant
mkdir -p gpurun_out
WL=$1; shift
for L in "$@"; do
  ORGPU_LIB=$PWD/build/liborgpu_$L.so python bench.py --workload $WL --steps 60 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$L', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])" | tee -a gpurun_out/sweep.log
done
