#!/bin/bash
# Round 2, GPU session C: frames read in place (span) - parity suite, bench with side legs, A/B against the gather.
export PYTHONPATH=$PWD; TAG=${1:-r2c}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log)
tail -25 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
RVAE_SPAN=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_nospan.json 2> gpurun_out/${TAG}_bench_nospan.err; echo "bench nospan rc=$?"
TAG=$TAG python - <<'P'
import json
import os
T = os.environ["TAG"]
for f in (T + "_bench", T + "_bench_nospan"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.3f M  e2e %.3f M" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6))
    for k in ("fp32_mode", "stream_4096", "widened_inference"):
        v = d.get(k, {})
        print("   ", k, v.get("error") or "%.3f M frames/s" % (v["value"] / 1e6), v.get("ms_per_step", v.get("ms_total")))
P
