# round 2, first GPU batch: full GPU suite, plastic bench, kernel variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_r2a.log
python -m pytest tests/test_ref_gpu_pin.py tests/test_qa_decks_gpu.py tests/test_balance_gpu.py -m gpu -q -s 2>&1 | grep -E "bending|cycle |passed|failed|Error|error" | tail -80 > gpurun_out/pytest_r2a_pin.log
for WL in c2_plate_qeph_1m c2_plate_qeph_1m_elastic c5_brick_slab_2m; do
  python bench.py --workload $WL --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/bench_r2a_${WL}.err | tail -1 > gpurun_out/bench_r2a_${WL}.json
done
rm -f gpurun_out/sweep.log
bash scripts/gpu_sweep.sh c2_plate_qeph_1m fixipla unroll5 fix_unroll
