mkdir -p gpurun_out
ORGPU_LIB=$PWD/build/liborgpu_m20.so python -m pytest tests/test_shell_gpu.py tests/test_full_size_gpu.py tests/test_sh3n_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_r2k.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest_r2k.log | cut -c1-300 | head -10
bash scripts/gpu_r2j.sh "$@"
