mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_r2d.log
tail -4 gpurun_out/pytest_r2d.log
for WL in c2_plate_qeph_1m c2_plate_qeph_1m_elastic c5_brick_slab_2m c1_taylor_bar; do
  python bench.py --workload $WL --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/bench_r2d_${WL}.err | tail -1 > gpurun_out/bench_r2d_${WL}.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_r2d_${WL}.json')); print('$WL value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['config'].get('plastic_fraction'), 'e2e %.4g'%d['e2e']['value'], d['e2e'].get('pcie'))"
done
for L in "" ns47 ns111; do
  if [ -z "$L" ]; then python scripts/many_sg_bench.py 8192 64 16 4 | tail -1 | sed 's/^/ns15 /'; else ORGPU_LIB=$PWD/build/liborgpu_$L.so python scripts/many_sg_bench.py 8192 64 16 4 | tail -1 | sed "s/^/$L /"; fi
done | tee gpurun_out/many_sg_r2d.log
