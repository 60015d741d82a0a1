mkdir -p gpurun_out
cp oclradixsort_b200/libb200rs.so /tmp/lib_orig.so
for g in 8 4 16 8; do
  cp tools/_build/lib_g$g.so oclradixsort_b200/libb200rs.so
  echo "== LB_GROUP=$g" >> gpurun_out/s40_lbgroup.txt
  timeout 300 python tools/sweep.py 28 keys=38 pairs=21 scan= 2>&1 | grep -v Warn >> gpurun_out/s40_lbgroup.txt
done
cp /tmp/lib_orig.so oclradixsort_b200/libb200rs.so
