mkdir -p gpurun_out
for V in psync ps3 ps5 ps6 ps15; do
  if [ $V = base ]; then unset ORGPU_LIB; else export ORGPU_LIB=$PWD/build/liborgpu_$V.so; fi
  python bench.py --steps 400 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$V c2 ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])"
done
