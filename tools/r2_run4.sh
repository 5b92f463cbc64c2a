set -x
O=gpurun_out/r2d
mkdir -p $O
./tools/debug/tma_test > $O/tma_test.log 2>&1; cat $O/tma_test.log
timeout 200 compute-sanitizer --tool memcheck ./tools/debug/tma_test > $O/tma_test_memcheck.log 2>&1; tail -30 $O/tma_test_memcheck.log
DETEXB200_TMA_STORE=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py -m gpu -x -q 2>&1 | tail -8 > $O/pytest_pipeline_notma.log; tail -4 $O/pytest_pipeline_notma.log
DETEXB200_TMA_STORE=0 timeout 300 python tools/sweep.py --formats BPTC,BPTC_FLOAT,BPTC_SIGNED_FLOAT --modes both > $O/sweep_pipeline_notma.log 2>&1; cat $O/sweep_pipeline_notma.log
DETEXB200_TMA_STORE=0 timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 > $O/sweep_bc7_16m_pipeline_notma.log 2>&1; cat $O/sweep_bc7_16m_pipeline_notma.log
timeout 300 python tools/debug/tma_perm.py > $O/tma_perm.log 2>&1; tail -40 $O/tma_perm.log
timeout 300 compute-sanitizer --tool memcheck python tools/debug/tma_perm.py > $O/tma_perm_memcheck.log 2>&1; grep -v "^=========     at\|^=========     by\|Host Frame\|^=========         " $O/tma_perm_memcheck.log | head -60
