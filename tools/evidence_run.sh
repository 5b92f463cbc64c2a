# One gpurun call that produces the round-2 evidence under gpurun_out/final/: GPU tests, the default bench (all legs) and the
# reference arm, sweeps, ncu --set full of the bench kernels (summaries + DRAM traffic recorded with the build id), the
# launch list of a bench run, sanitizer passes.   gpurun --timeout 2400 -- "bash tools/evidence_run.sh"
set -x
O=gpurun_out/final
mkdir -p gpurun_out/final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/nvidia_smi.txt
nproc > $O/nproc.txt
python -c "import detex_b200 as D; print(D.load().detexb200Version().decode())" > $O/version.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
cap() {   # name, traffic key ("" = none), kernel regex, skip, command...
	local name=$1 key=$2 rx=$3 skip=$4; shift 4
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/final/ncu_$name "$@" > $O/ncu_$name.log 2>&1
	python tools/ncu_summary.py gpurun_out/final/ncu_$name.ncu-rep $O/ncu_$name.txt "ncu_$name, round 2 evidence run, $(cat $O/version.txt)" > /dev/null 2>&1
	[ -n "$key" ] && python tools/ncu_traffic.py gpurun_out/final/ncu_$name.ncu-rep "$key" >> $O/ncu_traffic.log 2>&1
	rm -f gpurun_out/final/ncu_$name.ncu-rep
}
cap bc1_bench "decode_linear_kernel<BC1>" decode_linear_kernel 4 python bench.py --workload bc1 --no-extra --steps 3 --warmup 3 --no-cpu
cap bc7_bench "decode_sorted_kernel<BPTC,linear>" decode_sorted 4 python bench.py --workload bc7 --no-extra --steps 3 --warmup 3 --no-cpu
cap bc6h_bench "decode_sorted_kernel<BPTC_FLOAT,linear>" decode_sorted 4 python bench.py --workload bc6h --no-extra --steps 3 --warmup 3 --no-cpu
cap bc1_1e8_bench "decode_tiled_kernel<BC1>" decode_tiled_kernel 4 python bench.py --workload bc1_1e8 --no-extra --steps 3 --warmup 3 --no-cpu
cap mixed_bench "decode_tiled_kernel<EAC_RG11>" decode_tiled_kernel 10 python bench.py --workload mixed --no-extra --steps 3 --warmup 3 --no-cpu
cap bc6h_signed "" decode_sorted 4 python tools/sweep.py --formats BPTC_SIGNED_FLOAT --modes M --steps 3
cap etc2_tiled "" decode_tiled_kernel 4 python tools/sweep.py --formats ETC2 --steps 3
cap bc7_tiled "" decode_sorted 10 python tools/sweep.py --formats BPTC --modes M --blocks 16777216 --steps 3
cp profiles/ncu_traffic.json $O/ncu_traffic.json
timeout 900 python bench.py > $O/bench_all.json 2> $O/bench_all.err; tail -c 400 $O/bench_all.json
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --workload bc7u --no-extra > $O/bench_bc7u.json 2> $O/bench_bc7u.err
timeout 600 python tools/sweep.py --modes both 2>&1 | tee $O/sweep_all_formats.log | tail -3
timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 > $O/sweep_bc7_16m.log 2>&1
timeout 300 python tools/sweep.py --formats BPTC --modes M --bc7-single-modes 0,1,2,3,4,5,6,7 --blocks 8388608 > $O/sweep_bc7_single_modes.log 2>&1
timeout 300 python tools/sweep.py --formats ETC2,ETC2_PUNCHTHROUGH,ETC2_EAC --blocks 33554432 --steps 10 > $O/sweep_etc2_33m.log 2>&1
timeout 300 python tools/convert_bench.py 2>&1 | tee $O/convert_bench.log | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > $O/launches_bench.log 2>&1
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> $O/sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_check.py 2>&1 | grep -v "^=========     \|Host Frame\|^$" | tail -6 >> $O/sanitizer.txt
done
ls -la $O
