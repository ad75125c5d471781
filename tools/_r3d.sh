O=gpurun_out; T=r3d
timeout 600 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
for w in ppa cifar mnist; do timeout 200 python tools/kprof.py $w 8 > $O/${T}_kprof_$w.txt 2> $O/${T}_kprof_$w.err; done
PHC_B200_LIB=$PWD/phc_gnn_b200/variants/libphc_b200_prof.so TC_PROF=1 timeout 120 python tools/tc_bench.py 4 500 15616 1 10 > $O/${T}_tcprof_prof.txt 2>&1
tail -3 $O/${T}_pytest.log
