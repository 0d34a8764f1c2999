#!/bin/bash
# ncu capture of one steady-state nprior_kernel launch (the second-to-last launch of the job is a
# full chunk of 256 sweeps after the start-up transient).  Run from the repo root under gpurun.
set -e
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name nprior_kernel --csv \
    --log-file gpurun_out/np_launches.csv python tools/nprior_profile.py 128 600 0 > /dev/null 2>&1
n=$(grep -c nprior_kernel gpurun_out/np_launches.csv)
skip=$((n - 2))
echo "nprior_kernel launches: $n, capturing launch $skip"
ncu --set full --clock-control none --import-source on --kernel-name nprior_kernel --launch-skip $skip \
    --launch-count 1 -o gpurun_out/nprior_steady -f python tools/nprior_profile.py 128 600 0 > gpurun_out/ncu_nprior.log 2>&1
tail -1 gpurun_out/ncu_nprior.log
