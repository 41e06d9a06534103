#!/bin/bash
# Round 2, final evidence refresh (final code): ncu launch list of the bench command, HBM-kernel DRAM traffic, step timeline.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-sustained --no-graph > gpurun_out/r2x_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'adam_kernel|frame_gather|latent_bwd|randn_kernel|overlap_add|loss_fwd|loss_bwd|reparam_kernel|tanh_bwd|split_bf16|colsum' \
  --csv --log-file gpurun_out/r2x_hbm_kernels.csv python tools/ncu_hbm_kernels.py --ncu > gpurun_out/r2x_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
timeout 300 python tools/ncu_hbm_kernels.py > gpurun_out/r2x_hbm_events.log 2>&1; tail -14 gpurun_out/r2x_hbm_events.log | cut -c1-130
STEP_PIPE=1 TRACE_LAUNCH=0,1,2,3,4,5,6,7,8 TRACE_EPI_DETAIL=1 timeout 300 python tools/trace_step.py > gpurun_out/r2x_step_timeline.txt 2>&1; echo "trace rc=$?"; grep -A32 "merged timeline" gpurun_out/r2x_step_timeline.txt | tail -16 | cut -c1-120
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r2x_bench.json"))
print("value %.3f M  ms/step %.4f  e2e %.3f M  sustained %.3f M  roofline %.3f" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["sustained"]["value"] / 1e6, d["roofline"]["frac"]))
for k in ("fp32_mode", "stream_4096", "widened_inference"):
    v = d.get(k, {})
    print("   ", k, v.get("error") or "%.3f M frames/s" % (v["value"] / 1e6))
P
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "adam" 2>&1 | tail -2
