#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --blocks 15 > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2j_$name.json'))
    print('%-28s value %.3f M  ms/step %.4f  e2e %.3f M  first blocks %s' % ('$name', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['timing']['block_ms'][:3]))
except Exception as e:
    print('$name', 'FAILED', e)"
}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_train_step or cuda_graph or prefetched or full_size_gradient" 2>&1 | tail -2
for rep in 1 2; do
run lpt_$rep RVAE_CHAIN_ORDER=0
run spread_$rep RVAE_CHAIN_ORDER=1
run spread_ws4_$rep RVAE_CHAIN_ORDER=1 RVAE_WGRAD_SPLITS=4
run spread_ws8_$rep RVAE_CHAIN_ORDER=1 RVAE_WGRAD_SPLITS=8
run lpt_ws4_$rep RVAE_CHAIN_ORDER=0 RVAE_WGRAD_SPLITS=4
done
RVAE_CHAIN_ORDER=1 STEP_PIPE=1 TRACE_LAUNCH=4,7 TRACE_PAIRS=0,40 timeout 300 python tools/trace_step.py > gpurun_out/r2j_step_timeline.txt 2>&1; grep -A40 "merged timeline" gpurun_out/r2j_step_timeline.txt | tail -16 | cut -c1-160; grep -A12 "roles of launch 7" gpurun_out/r2j_step_timeline.txt | cut -c1-200
