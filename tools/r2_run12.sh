O=gpurun_out/r2n
mkdir -p gpurun_out/r2n
for tool in memcheck racecheck synccheck; do
  echo "== $tool" | tee -a $O/sanitizer.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_check.py 2>&1 | grep -v "^=========     \|Host Frame\|^$" | tail -12 | tee -a $O/sanitizer.log
done
echo "== memcheck, TMA off (cooperative copy-out of the staged tiles)" | tee -a $O/sanitizer.log
DETEXB200_TMA_STORE=0 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_check.py 2>&1 | grep -v "^=========     \|Host Frame\|^$" | tail -6 | tee -a $O/sanitizer.log
echo "== racecheck, TMA off" | tee -a $O/sanitizer.log
DETEXB200_TMA_STORE=0 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_check.py 2>&1 | grep -v "^=========     \|Host Frame\|^$" | tail -6 | tee -a $O/sanitizer.log
