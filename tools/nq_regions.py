"""Region / stall breakdown of an ncu source-page CSV of rollout_nq_kernel (developer tool)."""
import collections, csv, subprocess, sys
csvf, src = sys.argv[1], sys.argv[2]
T = float(sys.argv[3]) if len(sys.argv) > 3 else 151552.0
marks = [("layout", "nq_layout(int Vp"), ("commit", "struct NqCommit"), ("append", "void nq_append_warp"), ("pop", "uint32_t nq_pop"),
         ("scan fn", "void nq_scan_arrivals"), ("emit fns", "struct NqEmit"), ("decl/setup", "rollout_nq_kernel(DevParams"),
         ("window start", "---- window start"), ("fresh/valid", "if (fresh) {"), ("tick head", "for (int k = k0 - 1"),
         ("arrivals", "for (;;)"), ("prefill", "if (first_round) {"), ("emit call", "nq_emit_before_match<THREADS>(em"),
         ("classify", "-- classify"), ("batch path", "-- the other clusters: warps"), ("solo generic", "---- one cluster on its own"),
         ("stats flush", "const unsigned wm"), ("post+scan", "---- after the match"), ("end", "---- window end")]
lines = open(src).read().split("\n")
b = []
for name, pat in marks:
    for i, l in enumerate(lines, 1):
        if pat in l:
            b.append((name, i)); break
b.append(("eof", 100000))
rows = list(csv.reader(open(csvf)))
hdr = None; cur = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] == "": continue
    d = dict(zip(hdr[4:], r[4:]))
    try: ins = int(d["Instructions Executed"]); smp = int(d["# Samples"])
    except Exception: continue
    out.append((cur, int(r[0]), ins, smp, d))
ti = sum(o[2] for o in out); ts = sum(o[3] for o in out)
acc = collections.OrderedDict((n, [0, 0, collections.Counter()]) for n, _ in b[:-1])
other = collections.Counter(); others = collections.Counter()
sk = [k for k in hdr[4:] if k.startswith("stall_") and "Not Issued" not in k]
for f, ln, ins, smp, d in out:
    if f.startswith("rollout_nq"):
        for (n, a), (_, e) in zip(b[:-1], b[1:]):
            if a <= ln < e:
                acc[n][0] += ins; acc[n][1] += smp
                for k in sk:
                    try: acc[n][2][k] += int(d[k])
                    except Exception: pass
                break
    else:
        other[f] += ins; others[f] += smp
for n, (i, s_, st) in acc.items():
    top = ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(1, s_)) for k, v in st.most_common(4))
    print("%-13s %5.1f%%i %5.1f%%s %6.0f i/tick | %s" % (n, 100 * i / ti, 100 * s_ / ts, i / T, top))
for n, i in other.most_common():
    print("%-30s %5.1f%%i %5.1f%%s %6.0f" % (n, 100 * i / ti, 100 * others[n] / ts, i / T))
print("instr/tick %.0f  samples %d" % (ti / T, ts))
