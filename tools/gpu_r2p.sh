#!/bin/bash
# NVLS all-reduce: hardware DP tests + A/B against the peer-load kernel at N ranks
export PYTHONPATH=$PWD
N=${1:-2}
mkdir -p gpurun_out
true
true
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus $N --steps 20 --warmup 5 --no-sustained > gpurun_out/r2p_n${N}_$name.json 2> gpurun_out/r2p_n${N}_$name.err
  python -c "
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/r2p_n${N}_$name.json') if l.startswith('{')][-1])
    print('%-22s value %.3f M  ms/step %.4f  e2e %.3f M  blocks %s | %s' % ('$name', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['timing']['block_ms'][:3], d['config']['exchange'][:40]))
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/r2p_n${N}_$name.err').read()[-1500:])"
}
for rep in 1 2; do
run nvls_$rep RVAE_DP_BACKEND=nvls
run p2p_$rep RVAE_DP_BACKEND=p2p
done
true
