# tuning builds of libphox.so under tune/<name>.so ; usage: bash scripts/build_variants.sh name "-DFLAG=1 ..." [name flags ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p tune
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Iinclude"
C=eic-opticks_b200/csrc
while [ $# -ge 2 ]; do
  n=$1; f=$2; shift 2
  ( $NV $f -Xptxas -v -c -o tune/$n.engine.o $C/phox_engine.cu 2> tune/$n.ptxas.log && \
    $NV -shared -o tune/$n.so tune/$n.engine.o $C/phox_bvh.o $C/phox_merge.o -lcudart_static -lpthread -ldl -lrt && rm tune/$n.engine.o && echo built $n ) &
done
wait
