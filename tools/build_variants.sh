#!/usr/bin/env bash
# Build variants of libphc_b200.so that differ only in phm_linear_tc.cu's compile-time switches (kernel experiments: one gpurun call
# compares them all; PHC_B200_LIB=<path> selects one, see _lib.load).  Objects of the other sources come from csrc/build/.
#   tools/build_variants.sh name1="-DFLAG=1 ..." name2="..." [base=GIT_REV]   (base=REV compiles that revision's phm_linear_tc.cu)
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/phc_gnn_b200/csrc
OUT=$ROOT/phc_gnn_b200/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p "$OUT" /tmp/phc_variants
python -c "import sys; sys.path.insert(0, '$ROOT'); from phc_gnn_b200 import _lib; _lib.build()"
OTHERS=$(ls "$CSRC"/build/*.o | grep -v phm_linear_tc.o)
pids=()
for spec in "$@"; do
  name=${spec%%=*}; val=${spec#*=}
  (
    src=$CSRC/phm_linear_tc.cu; defs=$val
    if [ "$name" = base ]; then
      git -C "$ROOT" show "$val:phc_gnn_b200/csrc/phm_linear_tc.cu" > /tmp/phc_variants/base_tc.cu
      src=/tmp/phc_variants/base_tc.cu; defs=""
    fi
    $NVCC $FLAGS $defs -I"$CSRC" -c "$src" -o /tmp/phc_variants/$name.o
    $NVCC $FLAGS -shared -o "$OUT/libphc_b200_$name.so" /tmp/phc_variants/$name.o $OTHERS
    echo "built $OUT/libphc_b200_$name.so ($defs)"
  ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
