mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_r2b.log
rm -f gpurun_out/sweep.log
for WL in c2_plate_qeph_1m c2_plate_qeph_1m_elastic; do
python bench.py --workload $WL --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/bench_r2b.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$WL fast', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['config']['plastic_fraction'])" | tee -a gpurun_out/sweep.log
ORGPU_NO_FAST=1 python bench.py --workload $WL --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/bench_r2b.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$WL nofast', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])" | tee -a gpurun_out/sweep.log
ORGPU_LIB=$PWD/build/liborgpu_minb2.so python bench.py --workload $WL --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/bench_r2b.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$WL minb2', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])" | tee -a gpurun_out/sweep.log
done
