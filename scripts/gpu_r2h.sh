mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r2h.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest_r2h.log | head -30
python scripts/many_sg_bench.py 8192 64 16 4 | tail -1 | tee gpurun_out/many_sg_r2h.log
ORGPU_TAB_MIN=100000 python scripts/many_sg_bench.py 8192 64 16 4 | tail -1 | sed 's/^/notab /' | tee -a gpurun_out/many_sg_r2h.log
