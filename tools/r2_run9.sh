set -x
O=gpurun_out/r2j
mkdir -p gpurun_out/r2j
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py tests/test_gpu_convert.py -m gpu -x -q 2>&1 | tail -12 > $O/pytest.log; tail -5 $O/pytest.log
for v in main t512 not3; do
  if [ $v = main ]; then R="python"; else R="python tools/ab_variant.py run $v"; fi
  echo "== $v" >> $O/sweep.log
  timeout 300 $R tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 >> $O/sweep.log 2>&1
  timeout 300 $R tools/sweep.py --formats BPTC --modes both >> $O/sweep.log 2>&1
  timeout 300 $R tools/sweep.py --formats BPTC --modes M --bc7-single-modes 0,1,2,3,4,5,6,7 --blocks 8388608 > $O/sweep_bc7_single_$v.log 2>&1
done
echo "== main, BC7 LINEAR direct" >> $O/sweep.log
DETEXB200_BC7_LINEAR_DIRECT=1 timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 >> $O/sweep.log 2>&1
DETEXB200_BC7_LINEAR_DIRECT=1 timeout 300 python tools/sweep.py --formats BPTC --modes both >> $O/sweep.log 2>&1
echo "== main, TMA off" >> $O/sweep.log
DETEXB200_TMA_STORE=0 timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 >> $O/sweep.log 2>&1
echo "== main BC6H" >> $O/sweep.log
timeout 300 python tools/sweep.py --formats BPTC_FLOAT,BPTC_SIGNED_FLOAT --modes both >> $O/sweep.log 2>&1
cat $O/sweep.log
DETEXB200_BC7_LINEAR_DIRECT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "BPTC or bptc or config_bc7" 2>&1 | tail -3
