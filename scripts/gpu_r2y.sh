mkdir -p gpurun_out
python -m pytest tests/test_shell_gpu.py tests/test_sh3n_gpu.py tests/test_qa_decks_gpu.py tests/test_restart_gpu.py tests/test_domains_gpu.py -m gpu -q 2>&1 | tail -4
ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:qeph_forces -s 210 -c 1 -o gpurun_out/prof_r2y_qeph_rates -f \
    python bench.py --workload c2_plate_qeph_1m_rates --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ncu_r2y.log 2>&1
tail -2 gpurun_out/ncu_r2y.log
