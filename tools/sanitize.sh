# tools/sanitize.sh -- compute-sanitizer over the reduced suite; logs land in gpurun_out/ (copy to profiles/ to keep)
mkdir -p gpurun_out
python tools/sanitizer_suite.py > gpurun_out/sanitizer_plain.log 2>&1 || { tail -20 gpurun_out/sanitizer_plain.log; exit 1; }
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 40 python tools/sanitizer_suite.py > gpurun_out/sanitizer_r2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|== |suite done" gpurun_out/sanitizer_r2_$tool.log | tail -20
done
