#!/bin/bash
# usage: profile_launches.sh <script.py> <outname>: per-launch device times (cold cache, serialised)
python $1 > gpurun_out/plain_$2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$2.csv python $1 > gpurun_out/ncu_$2.log 2>&1
cat gpurun_out/plain_$2.log
