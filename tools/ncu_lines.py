"""Per-source-line summary of an `ncu --page source --print-source cuda,sass --csv` export."""
import csv, sys
fn = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(fn)))
hdr = None; cur_file = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] == "": continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        ins = int(d["Instructions Executed"]); smp = int(d["# Samples"])
    except Exception: continue
    out.append((cur_file, int(r[0]), r[1].strip()[:90], ins, smp, d))
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print("total instr %.3e samples %d" % (ti, ts))
key = 4 if (len(sys.argv) > 3 and sys.argv[3] == "samples") else 3
for f, ln, src, ins, smp, d in sorted(out, key=lambda o: -o[key])[:top]:
    print("%-12s %4d %5.1f%%i %5.1f%%s  %s" % (f[:12], ln, 100 * ins / ti, 100 * smp / ts, src))
