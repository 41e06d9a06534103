#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for u in 1 2 4; do for w in 4 8 16; do
  echo "== gather unroll $u waves $w"; RVAE_GATHER_UNROLL=$u RVAE_GATHER_WAVES=$w timeout 120 python tools/ncu_hbm_kernels.py --only frame_gather 2>&1 | grep frame_gather
done; done 2>&1 | tee gpurun_out/r2n_gather_sweep.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-sustained --blocks 15 > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2n_$name.json'))
    print('%-28s value %.3f M  ms/step %.4f  e2e %.3f M  first blocks %s' % ('$name', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['timing']['block_ms'][:3]))
except Exception as e:
    print('$name', 'FAILED', e)"
}
for rep in 1 2; do
run u4_$rep RVAE_GATHER_UNROLL=4
run u1_$rep RVAE_GATHER_UNROLL=1
run u2_$rep RVAE_GATHER_UNROLL=2
run u1_w16_$rep RVAE_GATHER_UNROLL=1 RVAE_GATHER_WAVES=16
done
