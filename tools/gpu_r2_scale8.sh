#!/bin/bash
# Round 2, 8-GPU session: the bench at N = 8, 4, 2, 1 back to back on one box (what the driver's SCALE run does).
# (hardware DP tests at world 2 / 4 / 8: python -m pytest tests/test_gpu_fullsize.py -m gpu -k data_parallel)
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2t_scale_n$n.json 2> gpurun_out/r2t_scale_n$n.err; echo "n$n rc=$?"
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2t_scale_n1.json 2> gpurun_out/r2t_scale_n1.err; echo "n1 rc=$?"
python - <<'P'
import json
v = {}
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2t_scale_n{n}.json") if l.startswith("{")][-1])
        v[n] = d
        s = d.get("sustained") or {}
        print(n, "value %.3f M  ms/step %.4f  e2e %.3f M  sustained %.3f M | %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, s.get("value", 0) / 1e6, (d["config"].get("exchange") or "")[:50]))
    except Exception as e:
        print(n, "unreadable", e)
for n in (2, 4, 8):
    if 1 in v and n in v:
        print("efficiency N=%d: %.3f (sustained %.3f)" % (n, v[n]["value"] / (n * v[1]["value"]),
              (v[n].get("sustained") or {}).get("value", 0) / (n * (v[1].get("sustained") or {}).get("value", 1))))
P
