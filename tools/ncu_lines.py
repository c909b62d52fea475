"""Per-source-line summary of an ncu report (`--import-source on`, kernels built with -lineinfo):
share of executed warp instructions and of stall samples per CUDA line.
Usage: ncu_lines.py report.ncu-rep [min_pct]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    h = rows[hdr]
    ix, isamp = h.index("Instructions Executed"), h.index("# Samples")
    lines = [r for r in rows[hdr + 1:] if len(r) > ix and r[0].isdigit() and r[2] == "-"]
    tot = sum(float(r[ix] or 0) for r in lines)
    tots = sum(float(r[isamp] or 0) for r in lines)
    print(f"total warp instructions {tot:.4g}, samples {tots:.0f}")
    for r in lines:
        a, s = float(r[ix] or 0) / tot * 100, float(r[isamp] or 0) / max(tots, 1) * 100
        if a >= minpct or s >= minpct:
            print(f"{r[0]:>5} {a:5.1f}% inst {s:5.1f}% samp | {r[1].strip()[:120]}")


if __name__ == "__main__":
    main()
