set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py -m gpu -x -q 2>&1 | tail -6 > $O/pytest_main.log; tail -3 $O/pytest_main.log
timeout 900 python tools/ab_variant.py run s256 -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py -m gpu -x -q 2>&1 | tail -6 > $O/pytest_s256.log; tail -3 $O/pytest_s256.log
for v in main s256 intpipe; do
  if [ $v = main ]; then R="python"; else R="python tools/ab_variant.py run $v"; fi
  timeout 300 $R tools/sweep.py --formats BPTC,BPTC_FLOAT,BPTC_SIGNED_FLOAT --modes both > $O/sweep_4m_$v.log 2>&1
  timeout 300 $R tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 > $O/sweep_bc7_16m_$v.log 2>&1
  echo "== $v"; cat $O/sweep_4m_$v.log $O/sweep_bc7_16m_$v.log
done
timeout 300 python tools/ab_variant.py run s256 tools/sweep.py --formats BPTC --modes M --bc7-single-modes 0,1,2,3,4,5,6,7 --blocks 8388608 > $O/sweep_bc7_single_s256.log 2>&1
timeout 300 python tools/convert_bench.py --only BPTC_FLOAT > $O/convert_bc6h_main.log 2>&1
timeout 300 python tools/ab_variant.py run s256 tools/convert_bench.py --only BPTC_FLOAT > $O/convert_bc6h_s256.log 2>&1
cap() {   # name, variant, kernel regex, sweep args...
	local name=$1 var=$2 rx=$3; shift 3
	if [ $var = main ]; then R="python"; else R="python tools/ab_variant.py run $var"; fi
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 -f -o $O/ncu_$name $R tools/sweep.py "$@" --steps 3 > $O/ncu_$name.log 2>&1
	python tools/ncu_summary.py $O/ncu_$name.ncu-rep $O/ncu_$name.txt "$name, round 2" > /dev/null 2>&1
	ncu -i $O/ncu_$name.ncu-rep --page source --csv > $O/ncu_${name}_source.csv 2>/dev/null
	rm -f $O/ncu_$name.ncu-rep
}
cap bc7_m_s256 s256 decode_sorted --formats BPTC --modes M --blocks 16777216
cap bc6h_signed_fp main decode_sorted --formats BPTC_SIGNED_FLOAT --modes M
cap bc6h_unsigned_fp main decode_sorted --formats BPTC_FLOAT --modes M
