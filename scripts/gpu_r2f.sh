mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_r2f.log
tail -6 gpurun_out/pytest_r2f.log
python -m pytest tests/test_ref_gpu_pin.py -m gpu -q -s 2>&1 | grep -E "kinematic|bending" | tail -30 > gpurun_out/pytest_r2f_pin.log
/usr/bin/time -v python bench.py 2>gpurun_out/bench_r2f.err | tail -1 > gpurun_out/bench_r2f.json
grep -E "Elapsed|Maximum resident" gpurun_out/bench_r2f.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r2f.json')); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['roofline'].get('elastic_state'), d['config'].get('plastic_fraction'), 'e2e %.4g'%d['e2e']['value'], d['cpu_baseline'], d['other_configs'])"
/usr/bin/time -v python bench.py --impl reference --steps 20 --warmup 3 2>gpurun_out/bench_r2f_ref.err | tail -1 > gpurun_out/bench_r2f_ref.json; grep -E "Elapsed" gpurun_out/bench_r2f_ref.err; cat gpurun_out/bench_r2f_ref.json | cut -c1-300
