#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
STEP_PIPE=1 TRACE_LAUNCH=1,2,3,4,5,6,7 TRACE_PAIRS=0,1,40,41,100,101,146,147 TRACE_EPI_DETAIL=1 timeout 300 python tools/trace_step.py > gpurun_out/r2f_roles.txt 2>&1; echo rc=$?
grep -n "roles of launch" gpurun_out/r2f_roles.txt
