mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
for L in "$@"; do
  ORGPU_LIB=$PWD/build/liborgpu_$L.so python bench.py --workload c2_plate_qeph_1m --steps 200 --warmup 20 --no-cpu-baseline --no-extras 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$L', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['roofline'].get('elastic_state',{}).get('avg_launch_ms'))" | tee -a gpurun_out/sweep.log
done
