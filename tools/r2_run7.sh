set -x
O=gpurun_out/r2g
mkdir -p gpurun_out/r2g
for ns in 0 1500 3000 6000 12000 24000; do
  for v in main s256; do
    if [ $v = main ]; then R="python"; else R="python tools/ab_variant.py run $v"; fi
    echo "== stagger $ns $v" >> $O/stagger.log
    DETEXB200_STAGGER_NS=$ns timeout 300 $R tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 >> $O/stagger.log 2>&1
    DETEXB200_STAGGER_NS=$ns timeout 300 $R tools/sweep.py --formats BPTC,BPTC_FLOAT,BPTC_SIGNED_FLOAT --modes both >> $O/stagger.log 2>&1
  done
done
cat $O/stagger.log
DETEXB200_STAGGER_NS=6000 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "BPTC or bptc or config or forced or partition" 2>&1 | tail -4
export DETEXB200_STAGGER_NS=6000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_sorted -s 4 -c 1 -f -o gpurun_out/r2g/ncu_bc7_m_stagger python tools/sweep.py --formats BPTC --modes M --blocks 16777216 --steps 3 > $O/ncu_bc7_m_stagger.log 2>&1
python tools/ncu_summary.py gpurun_out/r2g/ncu_bc7_m_stagger.ncu-rep $O/ncu_bc7_m_stagger.txt "bc7 M linear, stagger 6000" > /dev/null 2>&1
ncu -i gpurun_out/r2g/ncu_bc7_m_stagger.ncu-rep --page source --csv > $O/ncu_bc7_m_stagger_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_sorted -s 10 -c 1 -f -o gpurun_out/r2g/ncu_bc7_m_tiled python tools/sweep.py --formats BPTC --modes M --blocks 16777216 --steps 3 > $O/ncu_bc7_m_tiled.log 2>&1
python tools/ncu_summary.py gpurun_out/r2g/ncu_bc7_m_tiled.ncu-rep $O/ncu_bc7_m_tiled.txt "bc7 M tiled, stagger 6000" > /dev/null 2>&1
ncu -i gpurun_out/r2g/ncu_bc7_m_tiled.ncu-rep --page source --csv > $O/ncu_bc7_m_tiled_source.csv 2>/dev/null
rm -f gpurun_out/r2g/ncu_bc7_m_stagger.ncu-rep gpurun_out/r2g/ncu_bc7_m_tiled.ncu-rep
