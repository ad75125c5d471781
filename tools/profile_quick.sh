#!/usr/bin/env bash
# Reduced ncu evidence of the ppa step (one GPU, ~2 min): launch list + full captures of the dominant kernels; in-situ CUPTI times.
# usage: tools/profile_quick.sh TAG      -> gpurun_out/TAG_*.txt
set -u
R=${1:-r02b}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${R}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${R}_pytest.log
timeout 200 python tools/kprof.py ppa 8 > $O/${R}_kprof_ppa.txt 2> $O/${R}_kprof_ppa.err
timeout 200 python tools/kprof.py hiv 8 > $O/${R}_kprof_hiv.txt 2> $O/${R}_kprof_hiv.err
TC_PROF=1 timeout 120 python tools/tc_bench.py 4 500 15616 1 10 > $O/${R}_tcprof.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1100 --csv --log-file $O/${R}_launches_ppa.csv \
  python tools/step_time.py ppa 2 > $O/${R}_launches_ppa.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"phm_tc_mix_v3|phm_tc_dh_v2|conv_fwd_sums|aggregate_bwd_node|bn_apply_fwd|bn_bwd_reduce|bn_apply_bwd|bn_finalize" -s 60 -c 16 \
  -o /tmp/${R}_ppa_full -f python tools/step_time.py ppa 1 > $O/${R}_ppa_full.log 2>&1
python tools/ncu_summary.py launches $O/${R}_launches_ppa.csv > $O/${R}_launches_ppa.txt 2>&1
python tools/ncu_summary.py kernel /tmp/${R}_ppa_full.ncu-rep > $O/${R}_ppa_full.txt 2>&1
rm -f $O/${R}_launches_ppa.csv
tail -2 $O/${R}_pytest.log
