#!/bin/bash
# Round 2, GPU session K: parity suite, bench (PCM16-resident corpus A/B), ncu evidence for the round-2 code.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log)
tail -6 gpurun_out/r2k_pytest.log | cut -c1-300
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --blocks 15 > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2k_$name.json'))
    print('%-28s value %.3f M  ms/step %.4f  e2e %.3f M  first blocks %s' % ('$name', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['timing']['block_ms'][:3]))
except Exception as e:
    print('$name', 'FAILED', e)"
}
for rep in 1 2; do
run pcm16_$rep RVAE_PCM16_RESIDENT=1
run f32_$rep RVAE_PCM16_RESIDENT=0
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/r2k_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 11 -c 11 -o gpurun_out/r2k_gemms -f \
  python tools/ncu_gemms.py > gpurun_out/r2k_ncu_gemms.log 2>&1; echo "ncu gemms rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm -s 27 -c 9 -o gpurun_out/r2k_step -f \
  python tools/step_few.py 4 > gpurun_out/r2k_ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'adam_kernel|frame_gather|latent_bwd|randn_kernel|overlap_add|loss_fwd|loss_bwd|reparam_kernel|tanh_bwd|split_bf16|colsum' \
  --csv --log-file gpurun_out/r2k_hbm_kernels.csv python tools/ncu_hbm_kernels.py --ncu > gpurun_out/r2k_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
STEP_PIPE=1 TRACE_LAUNCH=0,1,2,3,4,5,6,7,8 TRACE_EPI_DETAIL=1 timeout 300 python tools/trace_step.py > gpurun_out/r2k_step_timeline.txt 2>&1; echo "trace rc=$?"
ls -la gpurun_out/r2k_*.ncu-rep
