mkdir -p gpurun_out
python -m pytest tests/test_full_size_gpu.py -m gpu -q -k "c5_slab_2m_bricks_phased" 2>&1 | grep -E "^E|assert|passed|failed" | head -20 > gpurun_out/pytest_r2g.log
cat gpurun_out/pytest_r2g.log
python -m pytest tests/test_ref_gpu_pin.py -m gpu -q 2>&1 | tail -3
S=$(date +%s); python bench.py 2>gpurun_out/bench_r2g.err | tail -1 > gpurun_out/bench_r2g.json; echo "bench default wall $(( $(date +%s) - S )) s"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2g.json')); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['roofline'].get('elastic_state'), d['config'].get('plastic_fraction'), 'e2e %.4g'%d['e2e']['value'], d['cpu_baseline']); print(json.dumps(d['other_configs']))"
S=$(date +%s); python bench.py --impl reference --steps 20 --warmup 3 2>gpurun_out/bench_r2g_ref.err | tail -1 > gpurun_out/bench_r2g_ref.json; echo "ref wall $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/bench_r2g_ref.json
