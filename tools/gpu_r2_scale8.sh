#!/bin/bash
# Round 2, 8-GPU session: hardware DP tests at world 2/4/8, then the bench at N = 8, 4, 2, 1 back to back (the
# driver's SCALE run), then one 8-rank step timeline.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "data_parallel" > gpurun_out/r2s_pytest_dp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest_dp.log)
tail -4 gpurun_out/r2s_pytest_dp.log | cut -c1-300
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2s_scale_n$n.json 2> gpurun_out/r2s_scale_n$n.err; echo "n$n rc=$?"
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2s_scale_n1.json 2> gpurun_out/r2s_scale_n1.err; echo "n1 rc=$?"
python - <<'P'
import json
v = {}
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2s_scale_n{n}.json") if l.startswith("{")][-1])
        v[n] = d
        s = d.get("sustained") or {}
        print(n, "value %.3f M  ms/step %.4f  e2e %.3f M  sustained %.3f M" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, s.get("value", 0) / 1e6))
    except Exception as e:
        print(n, "unreadable", e)
for n in (2, 4, 8):
    if 1 in v and n in v:
        print("efficiency N=%d: %.3f (sustained %.3f)" % (n, v[n]["value"] / (n * v[1]["value"]),
              (v[n].get("sustained") or {}).get("value", 0) / (n * (v[1].get("sustained") or {}).get("value", 1))))
P
STEP_PIPE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 tools/trace_step.py > gpurun_out/r2s_step_timeline_dp8.txt 2>&1; grep -A60 "merged timeline" gpurun_out/r2s_step_timeline_dp8.txt | tail -30 | cut -c1-200
