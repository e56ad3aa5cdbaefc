# compute-sanitizer passes over tools/sanitize.py (the default frame paths) and, for the bounce-queue counting sort
# and the canonical-node kernels, the same workload with the knobs that select them
set -o pipefail
out=gpurun_out/r2_sanitizer.txt; : > $out
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize.py" >> $out
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -E "sanitize workload|ERROR SUMMARY|RACECHECK SUMMARY|=========     at|Invalid|hazard|Uninit" | head -20 >> $out
done
echo "== MB200_SORT_BOUNCES=3 MB200_NODE_OCT=0 compute-sanitizer --tool memcheck python tools/sanitize.py" >> $out
MB200_SORT_BOUNCES=3 MB200_NODE_OCT=0 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | grep -E "sanitize workload|ERROR SUMMARY|=========     at|Invalid" | head -20 >> $out
echo "== MB200_SORT_BOUNCES=3 compute-sanitizer --tool racecheck python tools/sanitize.py" >> $out
MB200_SORT_BOUNCES=3 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize.py 2>&1 | grep -E "sanitize workload|RACECHECK SUMMARY|hazard" | head -20 >> $out
cat $out
