mkdir -p gpurun_out
python -m pytest tests/test_shell_gpu.py tests/test_sh3n_gpu.py tests/test_ref_gpu_pin.py tests/test_shell_gpu_abi.py tests/test_qa_decks_gpu.py tests/test_domains_gpu.py -m gpu -q 2>&1 | tail -4
for WL in tri_plate_1m_yielding bt_plate_1m_yielding; do
  for NF in 0 1; do
    if [ $NF = 1 ]; then export ORGPU_NO_FAST=1; else unset ORGPU_NO_FAST; fi
    python bench.py --workload $WL --steps 400 --no-cpu-baseline --no-extras 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$WL', 'generic' if $NF else 'three-pass', 'ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['config']['plastic_fraction'])"
  done
done
