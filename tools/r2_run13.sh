O=gpurun_out/r2o
mkdir -p gpurun_out/r2o
nvidia-smi --query-gpu=index,name --format=csv | tee $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_set or threads" 2>&1 | tail -4 | tee $O/pytest_devices.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 600 $O/bench_n2.json; tail -4 $O/bench_n2.err
timeout 600 python bench.py --impl reference --gpus 2 --steps 10 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python tools/e2e_devices.py --workloads bc1,bc7 2>&1 | tee $O/e2e_devices.log
cp gpurun_out/e2e_devices.json $O/ 2>/dev/null
