#!/usr/bin/env bash
# The switches are compiled in on request only: tools/build_variants.sh abl="-DPHC_TC_ABLATE_SWITCHES=1 -DPHC_TC_PROF=1" and run this with
# PHC_B200_LIB=$PWD/phc_gnn_b200/variants/libphc_b200_abl.so (the default library ignores PHC_TC_ABLATE).
M=${1:-15616}
for ab in 0 2 32 64 96 34; do
  PHC_TC_ABLATE=$ab TC_PROF=1 python tools/tc_bench.py 4 500 $M 1 10 2>&1 | sed "s/^/[prec=1 ablate=$ab] /"
done
