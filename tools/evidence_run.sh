# One gpurun call that produces the round evidence under gpurun_out/final/ (tests, bench of every workload,
# --impl reference, sweeps, BC7 per-mode A/B of both kernels, ncu --set full of the three headline kernels, launch
# list); every .ncu-rep is summarised on the box (tools/ncu_summary.py) and only the BC7 one travels back (gpurun_out/ is
# capped at 64 MiB); the summaries copied into profiles/ come from it:  gpurun --timeout 2400 -- "bash tools/evidence_run.sh"
set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/nvidia_smi.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_bc1.json 2> $O/bench_bc1.err; tail -c 600 $O/bench_bc1.json
timeout 600 python bench.py --workload bc7 > $O/bench_bc7.json 2> $O/bench_bc7.err; tail -c 300 $O/bench_bc7.json
timeout 600 python bench.py --workload bc6h > $O/bench_bc6h.json 2> $O/bench_bc6h.err
timeout 900 python bench.py --workload mixed > $O/bench_mixed.json 2> $O/bench_mixed.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python tools/sweep.py --modes both 2>&1 | tee $O/sweep_all_formats.log | tail -5
timeout 300 python tools/sweep.py --formats BPTC --modes M --bc7-single-modes 0,1,2,3,4,5,6,7 --blocks 8388608 > $O/sweep_bc7_single_modes_sorted.log 2>&1
DETEXB200_BC7_KERNEL=flow timeout 300 python tools/sweep.py --formats BPTC --modes both --bc7-single-modes 0,1,2,3,4,5,6,7 --blocks 8388608 > $O/sweep_bc7_single_modes_flow.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_sorted -s 3 -c 1 -f -o $O/ncu_bc7_bench python bench.py --workload bc7 --steps 3 --warmup 3 --no-cpu > $O/ncu_bc7_bench.log 2>&1
python tools/ncu_summary.py $O/ncu_bc7_bench.ncu-rep $O/ncu_bc7_bench.txt "ncu_bc7_bench, this round's evidence run" > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_sorted -s 3 -c 1 -f -o $O/ncu_bc6h_bench python bench.py --workload bc6h --steps 3 --warmup 3 --no-cpu > $O/ncu_bc6h_bench.log 2>&1
python tools/ncu_summary.py $O/ncu_bc6h_bench.ncu-rep $O/ncu_bc6h_bench.txt "ncu_bc6h_bench, this round's evidence run" > /dev/null 2>&1; rm -f $O/ncu_bc6h_bench.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_linear_kernel -s 3 -c 1 -f -o $O/ncu_bc1_bench python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_bc1_bench.log 2>&1
python tools/ncu_summary.py $O/ncu_bc1_bench.ncu-rep $O/ncu_bc1_bench.txt "ncu_bc1_bench, this round's evidence run" > /dev/null 2>&1; rm -f $O/ncu_bc1_bench.ncu-rep
timeout 300 python tools/convert_bench.py 2>&1 | tee $O/convert_bench.log | tail -3
DETEXB200_TONE_FUSION=0 timeout 300 python tools/convert_bench.py --only BPTC_FLOAT 2>&1 | grep -v native > $O/convert_bench_bc6h_two_pass.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_sorted -s 3 -c 1 -f -o $O/ncu_bc6h_rgbx8_tiled python tools/convert_bench.py --only BPTC_FLOAT --target RGBX8 --layout tiled --steps 2 > $O/ncu_bc6h_rgbx8_tiled.log 2>&1
python tools/ncu_summary.py $O/ncu_bc6h_rgbx8_tiled.ncu-rep $O/ncu_bc6h_rgbx8_tiled.txt "ncu_bc6h_rgbx8_tiled, this round's evidence run" > /dev/null 2>&1; rm -f $O/ncu_bc6h_rgbx8_tiled.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_sorted -s 3 -c 1 -f -o $O/ncu_bc6h_rgbx8_linear python tools/convert_bench.py --only BPTC_FLOAT --target RGBX8 --layout linear --steps 2 > $O/ncu_bc6h_rgbx8_linear.log 2>&1
python tools/ncu_summary.py $O/ncu_bc6h_rgbx8_linear.ncu-rep $O/ncu_bc6h_rgbx8_linear.txt "ncu_bc6h_rgbx8_linear, this round's evidence run" > /dev/null 2>&1; rm -f $O/ncu_bc6h_rgbx8_linear.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_bc1.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/launches_bench_bc1.log 2>&1
ls -la $O
