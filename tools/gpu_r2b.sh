#!/bin/bash
# Round 2, GPU session B: full parity suite, bench with the side legs, reference arm, sanitizer, ncu evidence.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log)
tail -5 gpurun_out/r2b_pytest.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"
timeout 120 python tools/sanitize_step.py > gpurun_out/r2b_sanitize_plain.log 2>&1; echo "sanitize plain rc=$?"; tail -3 gpurun_out/r2b_sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py > gpurun_out/r2b_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  tail -4 gpurun_out/r2b_sanitizer_$tool.log | cut -c1-300
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/r2b_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'adam_kernel|frame_gather|latent_bwd|randn_kernel|overlap_add|loss_fwd|loss_bwd|reparam_kernel|tanh_bwd|split_bf16|colsum' \
  --csv --log-file gpurun_out/r2b_hbm_kernels.csv python tools/ncu_hbm_kernels.py --ncu > gpurun_out/r2b_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
STEP_PIPE=1 timeout 300 python tools/trace_step.py > gpurun_out/r2b_step_timeline.txt 2>&1; tail -40 gpurun_out/r2b_step_timeline.txt
