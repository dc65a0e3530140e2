mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_r2ad.log; tail -2 gpurun_out/pytest_r2ad.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
S=$(date +%s); python bench.py 2>gpurun_out/bench_r2ad.err | tail -1 > gpurun_out/bench_r2ad.json; echo "bench default wall $(( $(date +%s) - S )) s"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2ad.json')); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'elastic %.3f'%d['roofline']['elastic_state']['frac'], d['config'].get('plastic_fraction'), 'e2e %.4g'%d['e2e']['value'], 'cpu %.4g'%d['cpu_baseline']['value'], d['clocks']); print({k:(round(v['ms_per_step'],4), round(v.get('roofline_frac',0),3)) for k,v in d['other_configs'].items()})"
S=$(date +%s); python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_r2ad_driver.json; echo "driver-like wall $(( $(date +%s) - S )) s"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2ad_driver.json')); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'], d['gpu_launches'])"
S=$(date +%s); python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_r2ad_ref.json; echo "ref wall $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/bench_r2ad_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 615 -c 45 --csv --log-file gpurun_out/launches_r2ad_qeph.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_r2ad.log 2>&1
ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:qeph_forces -s 210 -c 1 -f -o gpurun_out/prof_r2ad_qeph \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_r2ad.log 2>&1
ls -la gpurun_out/prof_r2ad_qeph.ncu-rep | cut -c20-80
python bench.py --workload c2_plate_qeph_1m_rates --steps 400 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > gpurun_out/bench_r2ad_rates.json
python -c "
import json
d=json.load(open('gpurun_out/bench_r2ad_rates.json')); print('rates value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])"
