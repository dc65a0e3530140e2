mkdir -p gpurun_out
rm -f gpurun_out/sweep.log
for L in nox nob noxb noxba; do
  ORGPU_LIB=$PWD/build/liborgpu_$L.so python bench.py --workload c2_plate_qeph_1m --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$L', 'ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])" | tee -a gpurun_out/sweep.log
done
bash scripts/gpu_ncu.sh r2_qeph qeph_forces --workload c2_plate_qeph_1m
python scripts/ncu_raw_summary.py gpurun_out/prof_r2_qeph.ncu-rep > gpurun_out/prof_r2_qeph_raw.txt 2>&1
ncu -i gpurun_out/prof_r2_qeph.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_src_summary.py > gpurun_out/prof_r2_qeph_src.txt 2>&1
