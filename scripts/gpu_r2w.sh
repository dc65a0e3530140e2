mkdir -p gpurun_out
python -m pytest tests/test_shell_gpu.py tests/test_qa_decks_gpu.py tests/test_restart_gpu.py tests/test_domains_gpu.py -m gpu -q 2>&1 | tail -4
for NF in 0 1; do
  if [ $NF = 1 ]; then export ORGPU_NO_FAST=1; else unset ORGPU_NO_FAST; fi
  python bench.py --workload c2_plate_qeph_1m_rates --steps 400 --no-cpu-baseline --no-extras 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('rates', 'generic' if $NF else 'three-pass', 'ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['config']['plastic_fraction'])"
done
