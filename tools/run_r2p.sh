mkdir -p gpurun_out
timeout 300 python tools/xp_probe.py 28 2 8 16 32 2>&1 | tail -4 | tee gpurun_out/r2p_xp_local.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'exchange_partition' -s 1 -c 1 -o /tmp/r2p_xp python tools/xp_probe.py 28 16 once > gpurun_out/r2p_ncu_xp.log 2>&1
ncu -i /tmp/r2p_xp.ncu-rep --page raw --csv > gpurun_out/r2p_raw_xp.csv 2>/dev/null
ncu -i /tmp/r2p_xp.ncu-rep --page source --csv > gpurun_out/r2p_source_xp.csv 2>/dev/null
ls -la gpurun_out/r2p_*
