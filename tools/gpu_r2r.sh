#!/bin/bash
# fc3 weight gradient deferred into stage 2's fused launch: parity suite + same-box A/B + timeline
export PYTHONPATH=$PWD
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log)
tail -4 gpurun_out/r2r_pytest.log | cut -c1-300; grep -n "^E " gpurun_out/r2r_pytest.log | head -5
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-sustained --blocks 15 > gpurun_out/r2r_$name.json 2> gpurun_out/r2r_$name.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2r_$name.json'))
    print('%-16s value %.3f M  ms/step %.4f  e2e %.3f M  first blocks %s' % ('$name', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['timing']['block_ms'][:3]))
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/r2r_$name.err').read()[-800:])"
}
for rep in 1 2; do
run merge_$rep RVAE_MERGE_B3W=1
run split_$rep RVAE_MERGE_B3W=0
done
STEP_PIPE=1 TRACE_LAUNCH=5,6 TRACE_PAIRS=0,40 timeout 300 python tools/trace_step.py > gpurun_out/r2r_step_timeline.txt 2>&1; grep -A32 "merged timeline" gpurun_out/r2r_step_timeline.txt | tail -14 | cut -c1-120
