O=gpurun_out/r2q
mkdir -p gpurun_out/r2q
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest.log
for v in main nosnake; do
  if [ $v = main ]; then R="python"; else R="python tools/ab_variant.py run $v"; fi
  echo "== $v" >> $O/sweep.log
  timeout 300 $R tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 >> $O/sweep.log 2>&1
  timeout 300 $R tools/sweep.py --formats BPTC,BPTC_FLOAT,BPTC_SIGNED_FLOAT --modes both >> $O/sweep.log 2>&1
done
cat $O/sweep.log
