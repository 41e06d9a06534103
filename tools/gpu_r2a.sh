#!/bin/bash
# Round 2, GPU session A: parity tests, the bench under the driver's flags and at 200 steps, HBM-kernel tuning sweep.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log)
tail -5 gpurun_out/r2a_pytest.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_driver.json 2> gpurun_out/r2a_bench_driver.err; echo "bench driver rc=$?"
cut -c1-300 gpurun_out/r2a_bench_driver.json
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2a_bench_200.json 2> gpurun_out/r2a_bench_200.err; echo "bench 200 rc=$?"
cut -c1-300 gpurun_out/r2a_bench_200.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_driver2.json 2> gpurun_out/r2a_bench_driver2.err; echo "bench driver2 rc=$?"
cut -c1-300 gpurun_out/r2a_bench_driver2.json
for u in 1 2 4; do for b in 2 4 8; do
  echo "== adam unroll $u blocks/SM $b"; RVAE_ADAM_UNROLL=$u RVAE_ADAM_BPS=$b timeout 120 python tools/ncu_hbm_kernels.py --only adam 2>&1 | grep adam
done; done 2>&1 | tee gpurun_out/r2a_adam_sweep.log
timeout 300 python tools/ncu_hbm_kernels.py 2>&1 | tee gpurun_out/r2a_hbm_events.log
STEP_PIPE=1 STEP_GRAPH=1 timeout 300 python tools/plan_perf.py 2>&1 | tee gpurun_out/r2a_plan_perf.log
STEP_PIPE=1 timeout 300 python tools/trace_step.py > gpurun_out/r2a_step_timeline.txt 2>&1; tail -30 gpurun_out/r2a_step_timeline.txt
