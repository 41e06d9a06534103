"""profiles/r2_sass_summary.txt: per kernel of librvae_b200.so, the SASS instruction count and the Blackwell-native
mnemonics (needs only cuobjdump, no GPU):   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import re
import subprocess
import sys
from pathlib import Path

so = Path(__file__).resolve().parents[1] / "rawaudiovae_kelsey_b200" / "librvae_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "LDGMC", "HMMA"]
rows = []
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)", txt, re.S):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = name.replace("rvae::", "").split("(")[0].replace("void ", "")
    body = m.group(2)
    ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", body)
    c = {k: 0 for k in COLS}
    for op in ins:
        if op.startswith("UTCHMMA"):
            c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        for k in ("LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "LDGMC"):
            if op.startswith(k):
                c[k] += 1
        if op.startswith("HMMA"):
            c["HMMA"] += 1
    rows.append((name, len(ins), c))
print("# SASS summary of rawaudiovae_kelsey_b200/librvae_b200.so (default build, RVAE_EXPERIMENTS=0); tools/sass_summary.py")
print("# UTCHMMA = tcgen05.mma (bf16), LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add,")
print("# UTCBAR = tcgen05.commit (mbarrier arrive), SYNCS = mbarrier ops, LDGMC = multimem.ld_reduce (NVLS in-switch reduction);")
print("# HMMA would be the legacy mma.sync path (none).\n")
print(f"{'kernel':100s} {'instr':>7s} " + " ".join(f"{k:>12s}" for k in COLS))
for name, n, c in sorted(rows):
    print(f"{name[:100]:100s} {n:7d} " + " ".join(f"{c[k]:12d}" for k in COLS))
tot = {k: sum(r[2][k] for r in rows) for k in COLS}
print(f"{'TOTAL':100s} {sum(r[1] for r in rows):7d} " + " ".join(f"{tot[k]:12d}" for k in COLS))
