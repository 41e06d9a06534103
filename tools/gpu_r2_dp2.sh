#!/bin/bash
# Round 2, GPU session E (2 GPUs): hardware data-parallel tests, N=1 and N=2 bench back to back, DP timeline.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "data_parallel or torchrun" > gpurun_out/r2l_pytest_dp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_dp.log)
tail -15 gpurun_out/r2l_pytest_dp.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2l_scale_n1.json 2> gpurun_out/r2l_scale_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_scale_n2.json 2> gpurun_out/r2l_scale_n2.err; echo "n2 rc=$?"
tail -3 gpurun_out/r2l_scale_n2.err
python - <<'P'
import json
v = {}
for n in (1, 2):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2l_scale_n{n}.json") if l.startswith("{")][-1])
        v[n] = d
        print(n, "value %.3f M  ms/step %.4f  e2e %.3f M" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6), d.get("dp_check"))
    except Exception as e:
        print(n, "unreadable", e)
if 1 in v and 2 in v:
    print("efficiency N=2: %.3f" % (v[2]["value"] / (2 * v[1]["value"])))
P
STEP_PIPE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/trace_step.py > gpurun_out/r2l_step_timeline_dp2.txt 2>&1; tail -45 gpurun_out/r2l_step_timeline_dp2.txt | cut -c1-200
