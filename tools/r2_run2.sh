# Round 2, GPU call 2: GPU test-suite on the new host path, the default bench (all legs), sweep of every format, and the
# ncu captures round 1 never took (signed BC6H, the ETC2 family).   gpurun --timeout 2400 -- "bash tools/r2_run2.sh"
set -x
O=gpurun_out/r2b
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/nvidia_smi.txt
nproc > $O/nproc.txt; free -g > $O/free.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
( time timeout 900 python bench.py ) > $O/bench_all.json 2> $O/bench_all.err; tail -c 1500 $O/bench_all.json; tail -5 $O/bench_all.err
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python tools/sweep.py --modes both 2>&1 | tee $O/sweep_all_formats.log | tail -4
timeout 300 python tools/sweep.py --formats BPTC --modes both --blocks 16777216 --steps 10 > $O/sweep_bc7_16m.log 2>&1
timeout 300 python tools/sweep.py --formats BPTC --modes M --bc7-single-modes 0,1,2,3,4,5,6,7 --blocks 8388608 > $O/sweep_bc7_single.log 2>&1
cap() {   # name, kernel regex, sweep args...
	local name=$1 rx=$2; shift 2
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 -f -o $O/ncu_$name python tools/sweep.py "$@" --steps 3 > $O/ncu_$name.log 2>&1
	python tools/ncu_summary.py $O/ncu_$name.ncu-rep $O/ncu_$name.txt "$name, round 2 (build $(python -c 'import detex_b200 as D; print(D.build_id())'))" > /dev/null 2>&1
	ncu -i $O/ncu_$name.ncu-rep --page source --csv > $O/ncu_${name}_source.csv 2>/dev/null
	rm -f $O/ncu_$name.ncu-rep
}
cap bc6h_signed_m decode_sorted --formats BPTC_SIGNED_FLOAT --modes M
cap bc6h_unsigned_m decode_sorted --formats BPTC_FLOAT --modes M
cap bc7_m decode_sorted --formats BPTC --modes M --blocks 16777216
cap etc2 decode_linear_kernel --formats ETC2
cap etc2_punch decode_linear_kernel --formats ETC2_PUNCHTHROUGH
cap etc2_eac decode_linear_kernel --formats ETC2_EAC
cap bc3 decode_linear_kernel --formats BC3
ls -la $O
