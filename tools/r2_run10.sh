set -x
O=gpurun_out/r2k
mkdir -p gpurun_out/r2k
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12 > $O/pytest.log; tail -5 $O/pytest.log
for v in main defer16; do
  if [ $v = main ]; then R="python"; else R="python tools/ab_variant.py run $v"; fi
  echo "== $v" >> $O/sweep.log
  timeout 300 $R tools/sweep.py --formats ETC1,ETC2,ETC2_PUNCHTHROUGH,ETC2_EAC >> $O/sweep.log 2>&1
done
echo "== main, BC7 LINEAR direct" >> $O/sweep.log
DETEXB200_BC7_LINEAR_DIRECT=1 timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 >> $O/sweep.log 2>&1
DETEXB200_BC7_LINEAR_DIRECT=1 timeout 300 python tools/sweep.py --formats BPTC --modes both >> $O/sweep.log 2>&1
echo "== main" >> $O/sweep.log
timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 >> $O/sweep.log 2>&1
timeout 300 python tools/sweep.py --formats BPTC --modes both >> $O/sweep.log 2>&1
cat $O/sweep.log
DETEXB200_BC7_LINEAR_DIRECT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "BPTC or bptc or config_bc7" 2>&1 | tail -3
( time timeout 900 python bench.py ) > $O/bench_all.json 2> $O/bench_all.err; tail -3 $O/bench_all.err
for f in etc2 etc2_punch; do
F=ETC2; [ $f = etc2_punch ] && F=ETC2_PUNCHTHROUGH
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_linear_kernel -s 4 -c 1 -f -o gpurun_out/r2k/ncu_$f python tools/sweep.py --formats $F --steps 3 > $O/ncu_$f.log 2>&1
python tools/ncu_summary.py gpurun_out/r2k/ncu_$f.ncu-rep $O/ncu_$f.txt "$f linear, prefetch" > /dev/null 2>&1
ncu -i gpurun_out/r2k/ncu_$f.ncu-rep --page source --csv > $O/ncu_${f}_source.csv 2>/dev/null
rm -f gpurun_out/r2k/ncu_$f.ncu-rep
done
