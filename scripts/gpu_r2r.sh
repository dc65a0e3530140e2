mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_r2r.log; tail -2 gpurun_out/pytest_r2r.log
python bench.py --steps 200 --no-extras 2>gpurun_out/bench_r2r.err | tail -1 > gpurun_out/bench_r2r.json
python -c "
import json
d=json.load(open('gpurun_out/bench_r2r.json')); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], 'frac %.3f'%d['roofline']['frac']); print(json.dumps(d['e2e'])); print(d['cpu_baseline'])"
tail -3 gpurun_out/bench_r2r.err
