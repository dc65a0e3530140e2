# usage (under gpurun --gpus 8): the round's 8-GPU record: C2 weak + C3 strong / C5 / C4 (other_configs), then the exchange A/B
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 2>gpurun_out/scale8_r2s.err | tail -1 > gpurun_out/scale8_r2s.json
tail -2 gpurun_out/scale8_r2s.err
ORGPU_NO_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 300 --no-extras 2>/dev/null | tail -1 > gpurun_out/scale8_r2s_serial.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 300 --no-extras 2>/dev/null | tail -1 > gpurun_out/scale8_r2s_overlap.json
python bench.py --gpus 1 --steps 300 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale1_r2s.json
python - <<'PY'
import json
for k in ("scale8_r2s", "scale8_r2s_serial", "scale8_r2s_overlap", "scale1_r2s"):
    try:
        d = json.load(open(f"gpurun_out/{k}.json")); print(k, "n", d["n_gpus"], "ms/step %.4f value %.4g" % (d["ms_per_step"], d["value"]), d.get("pon_check") and d["pon_check"]["bitwise_identical"], d["kernel_ms"])
        if d.get("other_configs"): print({a: (round(b["ms_per_step"], 4), "%.4g" % b["value"]) for a, b in d["other_configs"].items() if "value" in b})
    except Exception as e: print(k, "failed", e)
PY
