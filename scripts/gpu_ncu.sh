# usage: bash scripts/gpu_ncu.sh <tag> <kernel-regex> [bench args...]   (run under gpurun, one GPU)
mkdir -p gpurun_out
TAG=$1; KREG=$2; shift 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KREG} -s 3 -c 1 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
