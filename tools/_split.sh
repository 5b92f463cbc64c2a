O=gpurun_out/split; mkdir -p $O
for v in 0 1; do
  echo "== SPLIT=$v sweep" >> $O/ab.log
  DETEXB200_BC6H_SPLIT=$v python tools/sweep.py --formats BPTC_FLOAT,BPTC_SIGNED_FLOAT --modes both >> $O/ab.log 2>&1
  echo "== SPLIT=$v convert" >> $O/ab.log
  DETEXB200_BC6H_SPLIT=$v python tools/convert_bench.py --only BPTC_FLOAT,BPTC_SIGNED_FLOAT 2>&1 | grep decode >> $O/ab.log
done
DETEXB200_BC6H_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_convert.py tests/test_gpu_tone.py -m gpu -x -q -k "FLOAT or float or bc6h or BC6H or tone or convert or config" > $O/tests.log 2>&1
tail -4 $O/tests.log
cat $O/ab.log
