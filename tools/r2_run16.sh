O=gpurun_out/r2s
mkdir -p gpurun_out/r2s
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py tests/test_gpu_convert.py -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest.log
for v in main interleaved noscaled; do
  if [ $v = main ]; then R="python"; else R="python tools/ab_variant.py run $v"; fi
  echo "== $v" >> $O/sweep.log
  timeout 300 $R tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 >> $O/sweep.log 2>&1
  timeout 300 $R tools/sweep.py --formats BPTC,BPTC_FLOAT,BPTC_SIGNED_FLOAT --modes both >> $O/sweep.log 2>&1
done
cat $O/sweep.log
timeout 600 ncu --set full --clock-control none -k regex:decode_sorted -s 4 -c 1 -f -o gpurun_out/r2s/ncu_bc7 python tools/sweep.py --formats BPTC --modes M --blocks 16777216 --steps 3 > $O/ncu_bc7.log 2>&1
python tools/ncu_summary.py gpurun_out/r2s/ncu_bc7.ncu-rep $O/ncu_bc7.txt "bc7 M linear, blocked ids" > /dev/null 2>&1
rm -f gpurun_out/r2s/ncu_bc7.ncu-rep
grep -E "bank_conflicts|shared_st|shared_ld|duration" $O/ncu_bc7.txt
