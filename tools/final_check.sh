#!/usr/bin/env bash
# End-of-round check on one GPU: GPU test suite, smoke(), in-situ kernel times and the default bench line.
# usage: tools/final_check.sh TAG
set -u
T=${1:-final}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log; tail -3 $O/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/${T}_smoke.txt
timeout 200 python tools/kprof.py ppa 8 > $O/${T}_kprof_ppa.txt 2> /dev/null; head -3 $O/${T}_kprof_ppa.txt | grep -v Warn
tools/bench_round.sh $T ppa
