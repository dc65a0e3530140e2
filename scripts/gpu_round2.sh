# usage (under gpurun): bash scripts/gpu_round2.sh  -> GPU parity tests, then kernel-time sweep of the built variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r2.log; cat gpurun_out/pytest_r2.log
rm -f gpurun_out/sweep.log
for WL in c2_plate_qeph_1m c5_brick_slab_2m; do
  echo "== $WL" | tee -a gpurun_out/sweep.log
  bash scripts/gpu_sweep.sh $WL "$@"
done
