#!/usr/bin/env bash
# ncu evidence of one round (run under gpurun, ONE GPU): launch list of the ppa step, full captures of the dominant kernels at the
# ppa shape, of the softmax aggregation at the hiv shape and of the max aggregation at the ppa shape, and the prep kernels.
# usage: tools/profile_round.sh r02      -> gpurun_out/r02_*.{csv,ncu-rep,txt}
set -u
R=${1:-r02}
O=gpurun_out
# step_time.py: 8 warm-up steps, then 3 x STEPS timed steps.  A ppa step is ~270 launches: skip the warm-up, keep ~4 steps.
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1100 --csv --log-file $O/${R}_launches_ppa.csv \
  python tools/step_time.py ppa 2 > $O/${R}_launches_ppa.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"phm_tc_mix_v3|phm_tc_dh_v2|conv_fwd_sums|aggregate_bwd_node|bn_apply_fwd|bn_bwd_reduce|bn_apply_bwd|scan_kernel" -s 60 -c 16 \
  -o $O/${R}_ppa_full -f python tools/step_time.py ppa 1 > $O/${R}_ppa_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_fwd_kernel|aggregate_bwd" -s 8 -c 6 \
  -o $O/${R}_hiv_softmax_full -f python tools/step_time.py hiv 1 > $O/${R}_hiv_full.log 2>&1
STEP_AGGR=max timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_fwd_kernel|aggregate_bwd" -s 8 -c 6 \
  -o $O/${R}_ppa_max_full -f python tools/step_time.py ppa 1 > $O/${R}_ppa_max_full.log 2>&1
python tools/prep_bench.py > $O/${R}_prep.txt 2>&1
# gpurun_out/ travels back only below 64 MiB: keep the text summaries (what profiles/ commits), drop the reports
python tools/ncu_summary.py launches $O/${R}_launches_ppa.csv > $O/${R}_launches_ppa.txt 2>&1
for n in ppa_full hiv_softmax_full ppa_max_full; do
  python tools/ncu_summary.py kernel $O/${R}_$n.ncu-rep > $O/${R}_$n.txt 2>&1
  rm -f $O/${R}_$n.ncu-rep
done
rm -f $O/${R}_launches_ppa.csv
ls -la $O/${R}_*
