#!/bin/bash
# Round 2, final single-GPU validation: smoke, full parity suite, both bench arms exactly as the driver runs them.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log)
tail -4 gpurun_out/r2z_pytest.log | cut -c1-300
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r2z_bench.json"))
print("value %.3f M  ms/step %.4f  e2e %.3f M  sustained %.3f M  launches %d" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["sustained"]["value"] / 1e6, d["gpu_launches"]))
print("timing", {k: v for k, v in d["timing"].items() if "block_ms" not in k})
print("roofline", d["roofline"]["frac"], "whole", d["roofline_whole_step"]["frac"], d["roofline_whole_step"]["sustained_frac"], "clocks", d["clocks"])
for k in ("fp32_mode", "stream_4096", "widened_inference"):
    v = d.get(k, {})
    print("   ", k, v.get("error") or "%.3f M frames/s" % (v["value"] / 1e6), v.get("ms_per_step", v.get("ms_total")))
r = json.load(open("gpurun_out/r2z_bench_ref.json"))
print("reference arm %.1f frames/s, %.1f ms/step" % (r["value"], r["ms_per_step"]))
P
