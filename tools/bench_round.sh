#!/usr/bin/env bash
# N = 1 bench records of the BASELINE.json workloads (one JSON line each) -> gpurun_out/TAG_bench_<workload>_n1.json
# usage: tools/bench_round.sh TAG workload [workload ...]     (workload "zinc2" = zinc with --phm-dim 2, "ppa_bf16" = --precision bf16)
set -u
T=$1; shift
O=gpurun_out; mkdir -p $O
for w in "$@"; do
  case $w in
    zinc2) args="--workload zinc --phm-dim 2";;
    ppa_bf16) args="--workload ppa --precision bf16 --no-cpu-baseline";;
    *) args="--workload $w";;
  esac
  timeout 600 python bench.py $args > $O/${T}_bench_${w}_n1.json 2> $O/${T}_bench_${w}_n1.err
  echo "$w rc=$? $(python -c "import json,sys; d=json.loads(open('$O/${T}_bench_${w}_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'))" 2>&1 | tail -1)"
done
