mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_r2v.log; tail -2 gpurun_out/pytest_r2v.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
S=$(date +%s); python bench.py 2>gpurun_out/bench_r2v.err | tail -1 > gpurun_out/bench_r2v.json; echo "bench default wall $(( $(date +%s) - S )) s"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2v.json')); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'elastic %.3f'%d['roofline']['elastic_state']['frac'], d['config'].get('plastic_fraction'), 'e2e %.4g'%d['e2e']['value'], 'cpu %.4g'%d['cpu_baseline']['value'], d['clocks']); print({k:(round(v['ms_per_step'],4), round(v.get('roofline_frac',0),3)) for k,v in d['other_configs'].items()})"
S=$(date +%s); python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_r2v_driver.json; echo "driver-like wall $(( $(date +%s) - S )) s"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2v_driver.json')); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'], d['gpu_launches'])"
S=$(date +%s); python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_r2v_ref.json; echo "ref wall $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/bench_r2v_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 615 -c 45 --csv --log-file gpurun_out/launches_r2v_qeph.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_r2v.log 2>&1
ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:qeph_forces -s 210 -c 1 -f -o gpurun_out/prof_r2v_qeph \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_r2v.log 2>&1
ls -la gpurun_out/prof_r2v_qeph.ncu-rep | cut -c20-80
