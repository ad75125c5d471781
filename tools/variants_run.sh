#!/usr/bin/env bash
# One gpurun call: GPU test suite on the in-tree library, then the ppa step's in-situ kernel times (tools/kprof.py) for the in-tree
# library and every variant under phc_gnn_b200/variants/ (tools/build_variants.sh), parity of selected variants, role timers.
# usage: tools/variants_run.sh TAG [variants to parity-check ...]
set -u
TAG=${1:-v}; shift || true
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 200 python tools/kprof.py ppa 8 > $O/${TAG}_kprof_main.txt 2> $O/${TAG}_kprof_main.err
for so in phc_gnn_b200/variants/libphc_b200_*.so; do
  [ -e "$so" ] || continue
  n=$(basename $so .so); n=${n#libphc_b200_}
  PHC_B200_LIB=$PWD/$so timeout 200 python tools/kprof.py ppa 8 > $O/${TAG}_kprof_$n.txt 2> $O/${TAG}_kprof_$n.err
done
for n in "$@"; do
  PHC_B200_LIB=$PWD/phc_gnn_b200/variants/libphc_b200_$n.so timeout 300 python tools/tc_check.py > $O/${TAG}_check_$n.txt 2>&1
done
TC_PROF=1 timeout 120 python tools/tc_bench.py 4 500 15616 1 10 > $O/${TAG}_tcprof_main.txt 2>&1
if [ -e phc_gnn_b200/variants/libphc_b200_base.so ]; then
  PHC_B200_LIB=$PWD/phc_gnn_b200/variants/libphc_b200_base.so TC_PROF=1 timeout 120 python tools/tc_bench.py 4 500 15616 1 10 > $O/${TAG}_tcprof_base.txt 2>&1
fi
grep -h "ms/step\|phm_tc_mix_v3\|phm_tc_dh_v2\|bn_finalize" $O/${TAG}_kprof_*.txt
