# Round 2, GPU call 3: TMA store of the staged LINEAR tiles -- parity, then A/B against the cooperative copy-out.
set -x
O=gpurun_out/r2c
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py tests/test_gpu_convert.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
for v in 1 0; do
DETEXB200_TMA_STORE=$v timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 > $O/sweep_bc7_16m_tma$v.log 2>&1
DETEXB200_TMA_STORE=$v timeout 300 python tools/sweep.py --formats BPTC --modes both > $O/sweep_bc7_4m_tma$v.log 2>&1
DETEXB200_TMA_STORE=$v timeout 300 python tools/convert_bench.py --only BPTC_FLOAT --layout linear > $O/convert_bc6h_tma$v.log 2>&1
done
cat $O/sweep_bc7_16m_tma1.log $O/sweep_bc7_16m_tma0.log $O/sweep_bc7_4m_tma1.log $O/sweep_bc7_4m_tma0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_sorted -s 4 -c 1 -f -o $O/ncu_bc7_m_tma python tools/sweep.py --formats BPTC --modes M --blocks 16777216 --steps 3 > $O/ncu_bc7_m_tma.log 2>&1
python tools/ncu_summary.py $O/ncu_bc7_m_tma.ncu-rep $O/ncu_bc7_m_tma.txt "bc7 M linear, TMA store" > /dev/null 2>&1
ncu -i $O/ncu_bc7_m_tma.ncu-rep --page source --csv > $O/ncu_bc7_m_tma_source.csv 2>/dev/null
rm -f $O/*.ncu-rep
