#!/bin/bash
# Round 2, final 2-GPU validation: the whole -m gpu suite (multi-GPU cases included), bench at N = 1 and N = 2.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log)
tail -4 gpurun_out/r2w_pytest.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err; echo "n2 rc=$?"
python - <<'P'
import json
for n in (1, 2):
    d = json.loads([l for l in open(f"gpurun_out/r2w_bench_n{n}.json") if l.startswith("{")][-1])
    print(n, "value %.3f M  ms/step %.4f  e2e %.3f M  sustained %.3f M  launches %d" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["sustained"]["value"] / 1e6, d["gpu_launches"]), d["clocks"], d["sustained"]["clocks"])
P
