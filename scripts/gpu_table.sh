# usage (under gpurun): bash scripts/gpu_table.sh  -> one bench.py line per BASELINE workload at the default --steps/--warmup
mkdir -p gpurun_out
for WL in c1_taylor_bar c2_plate_qeph_1m c3_plate_qeph_4m c4_tube c5_brick_slab_2m tri_plate_1m; do
  python bench.py --workload $WL --no-cpu-baseline 2>gpurun_out/table_${WL}.err | tail -1 > gpurun_out/table_${WL}.json
  python -c "import json; d=json.load(open('gpurun_out/table_${WL}.json')); r=d['roofline']; print('$WL', 'ms/cycle %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'kernel_ms', {k:(round(v,4) if v else v) for k,v in d['kernel_ms'].items()}, 'force frac %.3f'%r['frac'], 'node GB/s %.0f'%r['node_kernel']['achieved'], 'cycle frac %.3f'%r['whole_cycle']['frac'], 'e2e %.4g'%d['e2e']['value'])"
done
