mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_r2m.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest_r2m.log | cut -c1-300 | head -10
S=$(date +%s); python bench.py 2>gpurun_out/bench_r2m.err | tail -1 > gpurun_out/bench_r2m.json; echo "bench default wall $(( $(date +%s) - S )) s"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2m.json')); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['roofline'].get('elastic_state'), d['config'].get('plastic_fraction'), 'e2e %.4g'%d['e2e']['value'], d['cpu_baseline']); print(json.dumps(d['other_configs']))"
S=$(date +%s); python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/bench_r2m_ref.err | tail -1 > gpurun_out/bench_r2m_ref.json; echo "ref wall $(( $(date +%s) - S )) s"; cut -c1-400 gpurun_out/bench_r2m_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 615 -c 45 --csv --log-file gpurun_out/launches_r2m_qeph.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_r2m.log 2>&1
tail -4 gpurun_out/launches_r2m_qeph.csv | cut -c1-300
ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:qeph_forces -s 210 -c 1 -f -o gpurun_out/prof_r2m_qeph \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_r2m.log 2>&1
tail -3 gpurun_out/ncu_full_r2m.log; ls -la gpurun_out/prof_r2m_qeph.ncu-rep
