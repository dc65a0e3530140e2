# usage: bash scripts/gpu_r2l.sh lib1 lib2 ...  -> C2 kernel times per variant in the plate's two regimes
mkdir -p gpurun_out; rm -f gpurun_out/sweep2.log
for L in "$@"; do
  for WS in "20 150" "200 400"; do
    set -- $WS
    ORGPU_LIB=$PWD/build/liborgpu_$L.so python bench.py --workload c2_plate_qeph_1m --warmup $1 --steps $2 --no-cpu-baseline --no-extras 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$L', 'W$1 K$2', 'ms/step %.4f'%d['ms_per_step'], 'shell %.4f node %.4f'%(d['kernel_ms']['shell_forces'], d['kernel_ms']['node']), 'frac %.3f'%d['roofline']['frac'], 'elastic %.4f'%d['roofline']['elastic_state']['avg_launch_ms'], d['config']['plastic_fraction']['last_cycle'])" | tee -a gpurun_out/sweep2.log
  done
done
