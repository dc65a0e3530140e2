mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r1b.log; cat gpurun_out/pytest_r1b.log
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_c2.log 2>gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.log; tail -5 gpurun_out/bench_c2.err
