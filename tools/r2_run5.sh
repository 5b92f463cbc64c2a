set -x
O=gpurun_out/r2e
mkdir -p $O
for k in 0 1 2 3 4 5 6 7; do ./tools/debug/tma_test $k; done > $O/tma_test.log 2>&1; cat $O/tma_test.log
timeout 300 python tools/debug/tma_perm.py > $O/tma_perm.log 2>&1; tail -12 $O/tma_perm.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py tests/test_gpu_convert.py -m gpu -x -q 2>&1 | tail -8 > $O/pytest_tma.log; tail -4 $O/pytest_tma.log
for v in 1 0; do
DETEXB200_TMA_STORE=$v timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 > $O/sweep_bc7_16m_tma$v.log 2>&1
DETEXB200_TMA_STORE=$v timeout 300 python tools/sweep.py --formats BPTC --modes both > $O/sweep_bc7_4m_tma$v.log 2>&1
DETEXB200_TMA_STORE=$v timeout 300 python tools/convert_bench.py --only BPTC_FLOAT --layout linear > $O/convert_bc6h_tma$v.log 2>&1
done
cat $O/sweep_bc7_16m_tma1.log $O/sweep_bc7_16m_tma0.log $O/sweep_bc7_4m_tma1.log $O/sweep_bc7_4m_tma0.log
timeout 300 python tools/sweep.py --formats BPTC --modes M --bc7-single-modes 0,1,2,3,4,5,6,7 --blocks 8388608 > $O/sweep_bc7_single.log 2>&1
cap() {   # name, kernel regex, sweep args...
	local name=$1 rx=$2; shift 2
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 -f -o $O/ncu_$name python tools/sweep.py "$@" --steps 3 > $O/ncu_$name.log 2>&1
	python tools/ncu_summary.py $O/ncu_$name.ncu-rep $O/ncu_$name.txt "$name, round 2 (build $(python -c 'import detex_b200 as D; print(D.build_id())'))" > /dev/null 2>&1
	ncu -i $O/ncu_$name.ncu-rep --page source --csv > $O/ncu_${name}_source.csv 2>/dev/null
	rm -f $O/ncu_$name.ncu-rep
}
cap bc7_m_tma decode_sorted --formats BPTC --modes M --blocks 16777216
cap bc6h_unsigned_m decode_sorted --formats BPTC_FLOAT --modes M
