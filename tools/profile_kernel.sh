#!/bin/bash
# usage: profile_kernel.sh <regex> <outname> <script.py> [skip] [count]
python $3 > gpurun_out/plain_$2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${4:-5} -c ${5:-2} -f -o gpurun_out/$2 python $3 > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/plain_$2.log
