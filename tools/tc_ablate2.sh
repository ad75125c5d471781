#!/usr/bin/env bash
M=${1:-15616}
for ab in 0 2 32 64 96 34; do
  PHC_TC_ABLATE=$ab TC_PROF=1 python tools/tc_bench.py 4 500 $M 1 10 2>&1 | sed "s/^/[prec=1 ablate=$ab] /"
done
