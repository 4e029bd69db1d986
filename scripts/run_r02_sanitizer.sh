# compute-sanitizer over the round-2 kernels at small shapes (scripts/sanitize_r02.py)
python scripts/sanitize_r02.py 2>&1 | tail -25
for part in attn rows train; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_r02.py $part > gpurun_out/san_mem_$part.log 2>&1
  echo "memcheck $part rc=$? $(grep -c PASS gpurun_out/san_mem_$part.log) pass $(grep 'ERROR SUMMARY' gpurun_out/san_mem_$part.log)"
done
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_r02.py attn > gpurun_out/san_race_attn.log 2>&1
echo "racecheck attn rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/san_race_attn.log)"
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_r02.py rows > gpurun_out/san_race_rows.log 2>&1
echo "racecheck rows rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/san_race_rows.log)"
