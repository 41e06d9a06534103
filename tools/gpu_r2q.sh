#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "per_kernel_timing or fused_train_step or full_size" 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2q_bench.err
python - <<'P'
import json
d = json.load(open("gpurun_out/r2q_bench.json"))
print("value %.3f M  ms/step %.4f  e2e %.3f M  sustained %.3f M" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["sustained"]["value"] / 1e6))
r = d["roofline"]; print({k: r[k] for k in ("achieved", "peak", "frac", "gemm_ms_per_step", "launches_per_step", "traffic")})
print({k: (round(v["us"], 1), v["tflops"] and round(v["tflops"])) for k, v in d["gemm_breakdown"].items() if isinstance(v, dict) and "us" in v})
print(d["gemm_breakdown"]["other_kernels_us_per_step"])
P
