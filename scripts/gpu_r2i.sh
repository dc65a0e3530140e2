mkdir -p gpurun_out
python -m pytest tests/test_shell_gpu.py tests/test_full_size_gpu.py tests/test_ref_gpu_pin.py tests/test_shell_gpu_abi.py tests/test_restart_gpu.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_r2i.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest_r2i.log | cut -c1-400 | head -30
python bench.py --no-extras --no-cpu-baseline 2>gpurun_out/bench_r2i.err | tail -1 > gpurun_out/bench_r2i.json
python -c "
import json
d=json.load(open('gpurun_out/bench_r2i.json')); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['roofline'].get('elastic_state'), d['config'].get('plastic_fraction'))"
