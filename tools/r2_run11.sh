O=gpurun_out/r2m
mkdir -p gpurun_out/r2m
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_api or device_set or threads or uniform or config_bc1 or multi_slab" 2>&1 | tail -5 | tee $O/pytest.log
python tools/e2e_pageable.py --host-bw --register 2>&1 | tail -1 | tee -a $O/pageable.log
for thr in 8 16 24; do for ring in 8 24 48; do for slab in 1024 4096; do
DETEXB200_POOL_RING=$ring DETEXB200_HOST_THREADS=$thr DETEXB200_POOL_SLAB_KB=$slab timeout 120 python tools/e2e_pageable.py 2>&1 | tail -1 | tee -a $O/pageable.log
done; done; done
DETEXB200_POOL_COPY=memcpy timeout 120 python tools/e2e_pageable.py 2>&1 | tail -1 | tee -a $O/pageable.log
