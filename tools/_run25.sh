mkdir -p gpurun_out
for pf in 0 64; do
echo "== PF=$pf" >> gpurun_out/s28_sorted.txt
B200RS_PF_TILES=$pf B200RS_TOOL_PROFILE=1 timeout 600 python tools/key_distributions.py 28 2>&1 | grep -E "sorted|uniform" >> gpurun_out/s28_sorted.txt
done
cd tools
B200RS_PF_TILES=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed -k regex:onesweep2 -s 4 -c 4 --csv --log-file ../gpurun_out/s28_ncu_sorted_pf0.csv python prof_dist.py 28 sorted > ../gpurun_out/s28_ncu.log 2>&1
B200RS_PF_TILES=64 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed -k regex:onesweep2 -s 4 -c 4 --csv --log-file ../gpurun_out/s28_ncu_sorted_pf64.csv python prof_dist.py 28 sorted >> ../gpurun_out/s28_ncu.log 2>&1
B200RS_PF_TILES=64 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed -k regex:onesweep2 -s 4 -c 4 --csv --log-file ../gpurun_out/s28_ncu_uniform_pf64.csv python prof_dist.py 28 uniform >> ../gpurun_out/s28_ncu.log 2>&1
