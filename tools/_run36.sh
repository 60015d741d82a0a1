mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> gpurun_out/s38_sanitizer.txt
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py 2>&1 | grep -v "^$" | tail -12 >> gpurun_out/s38_sanitizer.txt
done
