#!/bin/bash
python tools/probe_cycle.py > gpurun_out/plain_cycle.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cycle.csv python tools/probe_cycle.py > gpurun_out/ncu_cycle.log 2>&1
cat gpurun_out/plain_cycle.log
