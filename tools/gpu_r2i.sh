#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --blocks 15 > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/r2i_$name.json'))
print('%-28s value %.3f M  ms/step %.4f  e2e %.3f M  first blocks %s' % ('$name', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['timing']['block_ms'][:3]))"
}
for rep in 1 2; do
run nosplit_$rep RVAE_SPLIT_STAGE1=0
run o1_64_$rep RVAE_S1_ORDER=1 RVAE_S1_WGRAD_CTAS=64
run o0_64_$rep RVAE_S1_ORDER=0 RVAE_S1_WGRAD_CTAS=64
run o0_96_$rep RVAE_S1_ORDER=0 RVAE_S1_WGRAD_CTAS=96
run o0_64_bg80_$rep RVAE_S1_ORDER=0 RVAE_S1_WGRAD_CTAS=64 RVAE_ADAM_BG_BLOCKS=80
run o0_96_bg80_$rep RVAE_S1_ORDER=0 RVAE_S1_WGRAD_CTAS=96 RVAE_ADAM_BG_BLOCKS=80
run o0_64_bg160_$rep RVAE_S1_ORDER=0 RVAE_S1_WGRAD_CTAS=64 RVAE_ADAM_BG_BLOCKS=160
run nosplit_bg80_$rep RVAE_SPLIT_STAGE1=0 RVAE_ADAM_BG_BLOCKS=80
done
RVAE_S1_ORDER=0 RVAE_ADAM_BG_BLOCKS=80 STEP_PIPE=1 timeout 300 python tools/trace_step.py > gpurun_out/r2i_step_timeline.txt 2>&1; grep -A40 "merged timeline" gpurun_out/r2i_step_timeline.txt | tail -16 | cut -c1-160
