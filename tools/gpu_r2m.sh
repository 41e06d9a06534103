#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err
  python -c "
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/r2m_$name.json') if l.startswith('{')][-1])
    print('%-28s value %.3f M  ms/step %.4f  e2e %.3f M  blocks %s' % ('$name', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['timing']['block_ms'][:4]))
except Exception as e:
    print('$name', 'FAILED', e)"
}
for rep in 1 2; do
run default_$rep A=1
run nosplit_$rep RVAE_SPLIT_STAGE1=0
run lpt_$rep RVAE_CHAIN_ORDER=0
run nosplit_lpt_$rep RVAE_SPLIT_STAGE1=0 RVAE_CHAIN_ORDER=0
run split64_$rep RVAE_S1_WGRAD_CTAS=64
done
