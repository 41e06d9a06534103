#!/bin/bash
# Round 2, GPU session G: split backward stage 1 (latent kernel || fc3 weight gradient) - parity, A/B, timeline.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log)
tail -8 gpurun_out/r2g_pytest.log | cut -c1-300
for cfg in "1 64" "1 96" "1 32" "0 64"; do
  set -- $cfg
  echo "== RVAE_SPLIT_STAGE1=$1 RVAE_S1_WGRAD_CTAS=$2"
  RVAE_SPLIT_STAGE1=$1 RVAE_S1_WGRAD_CTAS=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --blocks 15 > gpurun_out/r2g_bench_$1_$2.json 2> gpurun_out/r2g_bench_$1_$2.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/r2g_bench_$1_$2.json'))
print('value %.3f M  ms/step %.4f  e2e %.3f M  first blocks %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['timing']['block_ms'][:4]))"
done
STEP_PIPE=1 TRACE_LAUNCH=5,6 timeout 300 python tools/trace_step.py > gpurun_out/r2g_step_timeline.txt 2>&1; grep -A40 "merged timeline" gpurun_out/r2g_step_timeline.txt | tail -32 | cut -c1-160
