O=gpurun_out/r2r
mkdir -p gpurun_out/r2r
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_convert.py tests/test_gpu_ingest.py -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest.log
for v in main planaronly; do
  if [ $v = main ]; then R="python"; else R="python tools/ab_variant.py run $v"; fi
  echo "== $v" >> $O/sweep.log
  timeout 300 $R tools/sweep.py --formats ETC1,ETC2,ETC2_PUNCHTHROUGH,ETC2_EAC >> $O/sweep.log 2>&1
  timeout 300 $R tools/sweep.py --formats ETC2,ETC2_PUNCHTHROUGH,ETC2_EAC --blocks 33554432 --steps 10 >> $O/sweep.log 2>&1
done
cat $O/sweep.log
for f in etc2 etc2_eac; do
F=ETC2; [ $f = etc2_eac ] && F=ETC2_EAC
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_linear_kernel -s 4 -c 1 -f -o gpurun_out/r2r/ncu_$f python tools/sweep.py --formats $F --steps 3 > $O/ncu_$f.log 2>&1
python tools/ncu_summary.py gpurun_out/r2r/ncu_$f.ncu-rep $O/ncu_$f.txt "$f linear, T/H/planar deferred" > /dev/null 2>&1
ncu -i gpurun_out/r2r/ncu_$f.ncu-rep --page source --csv > $O/ncu_${f}_source.csv 2>/dev/null
rm -f gpurun_out/r2r/ncu_$f.ncu-rep
done
