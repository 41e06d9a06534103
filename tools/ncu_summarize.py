"""Summarise an `ncu --set full` report of tools/ncu_gemms.py (the 11 GEMMs of a default.ini step, launched alone) into
the JSON bench.py reads for roofline.traffic:   python tools/ncu_summarize.py gpurun_out/X.ncu-rep profiles/r2_gemms_ncu.json
Needs only the ncu CLI (no GPU): it reads `ncu -i X.ncu-rep --page raw --csv`."""
import csv
import io
import json
import subprocess
import sys

NAMES = ["F1", "F2", "F3", "F4_out", "B4w", "B4d", "B3w", "B3d", "B2w", "B2d", "B1w"]
GFLOP = [34.36, 17.18, 8.59, 34.36, 34.36, 34.36, 8.59, 8.59, 17.18, 17.18, 34.36]
# --step: the report holds the 9 GEMM launches of ONE fused training step (tools/step_few.py): 2 of them are fused
# dgrad + weight-gradient launches
STEP_NAMES = ["F1", "F2", "F3", "F4_out", "B4d+B4w", "B3d", "B3w", "B2d+B2w", "B1w"]
STEP_GFLOP = [34.36, 17.18, 8.59, 34.36, 68.72, 8.59, 8.59, 34.36, 34.36]


def col(header, name):
    return header.index(name)


def main(rep, out, step=False):
    global NAMES, GFLOP
    if step:
        NAMES, GFLOP = STEP_NAMES, STEP_GFLOP
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    c = {
        "kernel": col(header, "Kernel Name"), "grid": col(header, "launch__grid_size"),
        "regs": col(header, "launch__registers_per_thread"),
        "us": col(header, "gpu__time_duration.sum"),
        "tens_el": col(header, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "tens_ac": col(header, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "dr": col(header, "dram__bytes_read.sum"), "dw": col(header, "dram__bytes_write.sum"),
        "xbar": col(header, "l1tex__m_xbar2l1tex_read_bytes.sum"),
    }
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "ns": 1e-3, "ms": 1e3, "s": 1e6}
    val = lambda r, k: float(r[c[k]].replace(",", "")) * scale.get(units[c[k]], 1.0)
    gemms = []
    for i, r in enumerate(data[-len(NAMES):]):
        us = val(r, "us")
        tel = float(r[c["tens_el"]])
        gemms.append({"gemm": NAMES[i], "kernel": r[c["kernel"]], "gflop": GFLOP[i], "us": us,
                      "tflops": GFLOP[i] / us * 1e3,
                      "tensor_active_pct_elapsed": tel, "tensor_active_pct_active": float(r[c["tens_ac"]]),
                      "dram_read_MB": val(r, "dr") / 1e6, "dram_write_MB": val(r, "dw") / 1e6,
                      "xbar2l1_read_MB": val(r, "xbar") / 1e6, "grid": float(r[c["grid"]]), "regs": float(r[c["regs"]])})
    tot_us = sum(g["us"] for g in gemms)
    summary = {
        "source": (f"ncu --set full --clock-control none -k regex:gemm -s 27 -c 9 python tools/step_few.py 4 ({rep}): the 9 GEMM "
                   "launches of one fused training step; serialised by ncu" if step else
                   f"ncu --set full --clock-control none -k regex:gemm_kernel -s 11 -c 11 python tools/ncu_gemms.py ({rep}); "
                   "per launch; cold cache, serialised"),
        "gemms": gemms,
        "sum_us": tot_us,
        "chain_tflops": sum(GFLOP) / tot_us * 1e3,
        "chain_weighted_tensor_pipe_active_pct_elapsed": sum(g["tensor_active_pct_elapsed"] * g["us"] for g in gemms) / tot_us,
        "dram_MB_per_step": sum(g["dram_read_MB"] + g["dram_write_MB"] for g in gemms),
        "dram_bytes_per_step": 1e6 * sum(g["dram_read_MB"] + g["dram_write_MB"] for g in gemms),
    }
    json.dump(summary, open(out, "w"), indent=1)
    for g in gemms:
        print(f"{g['gemm']:7s} {g['us']:6.1f} us  {g['tflops']:7.1f} TFLOP/s  tensor {g['tensor_active_pct_elapsed']:5.1f} % elapsed "
              f"{g['tensor_active_pct_active']:5.1f} % active  dram {g['dram_read_MB'] + g['dram_write_MB']:6.1f} MB  xbar {g['xbar2l1_read_MB']:6.1f} MB")
    print(f"sum {tot_us:.1f} us, chain {summary['chain_tflops']:.1f} TFLOP/s, chain-weighted tensor pipe "
          f"{summary['chain_weighted_tensor_pipe_active_pct_elapsed']:.1f} % of elapsed, DRAM {summary['dram_MB_per_step']:.0f} MB/step")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], "--step" in sys.argv)
