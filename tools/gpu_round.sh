#!/bin/bash
# One GPU session: parity tests, the bench (both arms), the ncu launch list of the bench command, and one full
# ncu capture of the 11 GEMMs of a training step. Outputs under gpurun_out/<tag>_*.
tag=${1:-r1}
export PYTHONPATH=$PWD
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log)
tail -3 gpurun_out/${tag}_pytest.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${tag}_bench.json
timeout 300 python tools/plan_perf.py 2>&1 | tee gpurun_out/${tag}_plan_perf.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 11 -c 11 -o gpurun_out/${tag}_gemms -f \
  python tools/ncu_gemms.py > gpurun_out/${tag}_ncu_gemms.log 2>&1; echo "ncu gemms rc=$?"
