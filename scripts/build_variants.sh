# builds liborgpu variants into build/ : usage  bash scripts/build_variants.sh NAME "-DFLAG=.. -DFLAG2=.." ...
set -e
mkdir -p build
cd openradioss_b200/csrc
while [ $# -gt 1 ]; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo ${FMAD:--fmad=false} -prec-div=true -prec-sqrt=true \
     -Xcompiler -fPIC -Xptxas -v $2 -shared -o ../../build/liborgpu_$1.so engine.cu 2> ../../build/ptxas_$1.log &
  shift 2
done
wait
