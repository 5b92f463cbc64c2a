# 8-GPU call: the default bench at N = 4 and N = 8 (all legs, per-rank parity), the reference arm, and ONE
# detexDecompressTextureLinear call spread over 1 / 2 / 4 / 8 GPUs.   gpurun --gpus 8 --timeout 1200 -- "bash tools/evidence_n8.sh"
O=gpurun_out/n8
mkdir -p gpurun_out/n8
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt; nproc > $O/nproc.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
for n in 8 4; do
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 ) > $O/bench_n$n.json 2> $O/bench_n$n.err; tail -c 300 $O/bench_n$n.json; tail -4 $O/bench_n$n.err
done
timeout 300 python bench.py --impl reference --gpus 8 --steps 10 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python tools/e2e_devices.py --workloads bc1,bc7 2>&1 | tee $O/e2e_devices.log
cp gpurun_out/e2e_devices.json $O/ 2>/dev/null
