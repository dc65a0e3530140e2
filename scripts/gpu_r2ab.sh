mkdir -p gpurun_out
run() { # variant workload
  if [ $1 = main ]; then unset ORGPU_LIB; else export ORGPU_LIB=$PWD/build/liborgpu_$1.so; fi
  python bench.py --workload $2 --steps 400 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1 $2 ms/step %.4f'%d['ms_per_step'], d['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])"
}
for W in bt_plate_1m_yielding tri_plate_1m_yielding c2_plate_qeph_1m_rates; do for V in nosync main; do run $V $W; done; done
for V in nosync b35 b67 b99; do run $V c5_brick_slab_2m; done
