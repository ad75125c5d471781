#!/usr/bin/env bash
# The switches are compiled in on request only: tools/build_variants.sh abl="-DPHC_TC_ABLATE_SWITCHES=1 -DPHC_TC_PROF=1" and run this with
# PHC_B200_LIB=$PWD/phc_gnn_b200/variants/libphc_b200_abl.so (the default library ignores PHC_TC_ABLATE).
# Ablation table of the tensor-core mix kernel at the ppa shape: which role bounds it (see MixParams::ablate).
# usage: tools/tc_ablate.sh [rows]   -> prints one line per (precision, ablation mask)
M=${1:-15616}
for prec in 1 2; do
  for ab in 0 1 2 4 8 16 3 5 6 9 12 17 24 7 15 31; do
    PHC_TC_ABLATE=$ab TC_PROF=1 python tools/tc_bench.py 4 500 $M $prec 10 2>&1 | sed "s/^/[prec=$prec ablate=$ab] /"
  done
done
