# usage (under gpurun --gpus 2): multi-GPU correctness + overlap A/B
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_domains_gpu.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_r2c_multi.log
tail -5 gpurun_out/pytest_r2c_multi.log
for WL in c2_plate_qeph_1m c5_brick_slab_2m; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --workload $WL --no-cpu-baseline 2>gpurun_out/r2c_${WL}.err | tail -1 > gpurun_out/r2c_overlap_${WL}.json
  ORGPU_NO_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 --workload $WL --no-cpu-baseline 2>gpurun_out/r2c_${WL}_no.err | tail -1 > gpurun_out/r2c_serial_${WL}.json
  python bench.py --gpus 1 --steps 200 --warmup 20 --workload $WL --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2c_single_${WL}.json
  python - <<PY
import json
for k in ("single","overlap","serial"):
    try:
        d=json.load(open("gpurun_out/r2c_%s_$WL.json"%k)); print("$WL",k,"ms/step %.4f value %.4g"%(d["ms_per_step"],d["value"]), d.get("pon_check"))
    except Exception as e: print("$WL",k,"failed",e)
PY
done
