#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 3 -o gpurun_out/prof_early python tools/ncu_target.py 64 1 > gpurun_out/ncu_early.log 2>&1
tail -2 gpurun_out/ncu_early.log
