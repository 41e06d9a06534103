#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for cfg in "0 64" "0 96" "0 128" "0 48" "1 64"; do
  set -- $cfg
  echo "== RVAE_S1_ORDER=$1 RVAE_S1_WGRAD_CTAS=$2"
  RVAE_S1_ORDER=$1 RVAE_S1_WGRAD_CTAS=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --blocks 15 > gpurun_out/r2h_bench_$1_$2.json 2> gpurun_out/r2h_bench_$1_$2.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/r2h_bench_$1_$2.json'))
print('value %.3f M  ms/step %.4f  e2e %.3f M  first blocks %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['timing']['block_ms'][:4]))"
done
RVAE_S1_ORDER=0 STEP_PIPE=1 TRACE_LAUNCH=5,6 timeout 300 python tools/trace_step.py > gpurun_out/r2h_step_timeline.txt 2>&1; grep -A40 "merged timeline" gpurun_out/r2h_step_timeline.txt | tail -16 | cut -c1-160
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_train_step or cuda_graph or prefetched" 2>&1 | tail -2
