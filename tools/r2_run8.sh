set -x
O=gpurun_out/r2i
mkdir -p gpurun_out/r2i
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py tests/test_gpu_convert.py -m gpu -x -q 2>&1 | tail -12 > $O/pytest_stream.log; tail -5 $O/pytest_stream.log
for v in stream classic; do
  echo "== $v" >> $O/sweep.log
  DETEXB200_SORT_KERNEL=$v timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 >> $O/sweep.log 2>&1
  DETEXB200_SORT_KERNEL=$v timeout 300 python tools/sweep.py --formats BPTC,BPTC_FLOAT,BPTC_SIGNED_FLOAT --modes both >> $O/sweep.log 2>&1
done
cat $O/sweep.log
timeout 300 python tools/sweep.py --formats BPTC --modes M --bc7-single-modes 0,1,2,3,4,5,6,7 --blocks 8388608 > $O/sweep_bc7_single.log 2>&1
timeout 300 python tools/convert_bench.py --only BPTC_FLOAT > $O/convert_bc6h.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 4 -c 1 -f -o gpurun_out/r2i/ncu_bc7_m_stream python tools/sweep.py --formats BPTC --modes M --blocks 16777216 --steps 3 > $O/ncu_bc7_m_stream.log 2>&1
python tools/ncu_summary.py gpurun_out/r2i/ncu_bc7_m_stream.ncu-rep $O/ncu_bc7_m_stream.txt "bc7 M linear, stream kernel" > /dev/null 2>&1
ncu -i gpurun_out/r2i/ncu_bc7_m_stream.ncu-rep --page source --csv > $O/ncu_bc7_m_stream_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 10 -c 1 -f -o gpurun_out/r2i/ncu_bc7_m_stream_tiled python tools/sweep.py --formats BPTC --modes M --blocks 16777216 --steps 3 > $O/ncu_bc7_m_stream_tiled.log 2>&1
python tools/ncu_summary.py gpurun_out/r2i/ncu_bc7_m_stream_tiled.ncu-rep $O/ncu_bc7_m_stream_tiled.txt "bc7 M tiled, stream kernel" > /dev/null 2>&1
ncu -i gpurun_out/r2i/ncu_bc7_m_stream_tiled.ncu-rep --page source --csv > $O/ncu_bc7_m_stream_tiled_source.csv 2>/dev/null
rm -f gpurun_out/r2i/ncu_bc7_m_stream.ncu-rep gpurun_out/r2i/ncu_bc7_m_stream_tiled.ncu-rep
