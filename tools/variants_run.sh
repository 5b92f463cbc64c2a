O=gpurun_out/variants
mkdir -p gpurun_out/variants
echo "== DETEXB200_DEVICES=all (every host-pointer texture call takes the pool route)" >> $O/pytest_variants.log
DETEXB200_DEVICES=all timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $O/pytest_variants.log
echo "== DETEXB200_TMA_STORE=0 (cooperative copy-out of the staged tiles)" >> $O/pytest_variants.log
DETEXB200_TMA_STORE=0 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $O/pytest_variants.log
echo "== DETEXB200_HOST_THREADS=2 DETEXB200_POOL_RING=2 DETEXB200_POOL_SLAB_KB=64 (tiny ring, tiny slabs)" >> $O/pytest_variants.log
DETEXB200_HOST_THREADS=2 DETEXB200_POOL_RING=2 DETEXB200_POOL_SLAB_KB=64 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3 >> $O/pytest_variants.log
echo "== DETEXB200_UNSORTED=1 (plain kernels for BC7 / BC6H)" >> $O/pytest_variants.log
DETEXB200_UNSORTED=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tone.py -m gpu -q 2>&1 | tail -3 >> $O/pytest_variants.log
cat $O/pytest_variants.log
