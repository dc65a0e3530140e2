mkdir -p gpurun_out
python -m pytest tests/test_shell_gpu.py tests/test_sh3n_gpu.py tests/test_qa_decks_gpu.py tests/test_restart_gpu.py tests/test_domains_gpu.py tests/test_full_size_gpu.py -m gpu -q 2>&1 | tail -6
