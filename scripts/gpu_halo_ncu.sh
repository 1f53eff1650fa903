#!/bin/bash
# ncu capture of the halo kernel (run under gpurun from the repo root); only CSV exports come back.
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -c 4 \
  -o /tmp/halo python scripts/gpu_halo_ncu_driver.py > gpurun_out/r02_ncu_halo.log 2>&1
tail -3 gpurun_out/r02_ncu_halo.log
ncu -i /tmp/halo.ncu-rep --page raw --csv > gpurun_out/r02_ncu_halo_raw.csv 2>/dev/null
ncu -i /tmp/halo.ncu-rep --page source --csv > gpurun_out/r02_ncu_halo_source.csv 2>/dev/null
ls -la gpurun_out/r02_ncu_halo_*.csv
