"""Per-line and per-phase view of an `ncu --set full --import-source on` capture of cmpc_solve_kernel.
  ncu -i REP --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_source_top.py src.csv [out.csv]
Writes the 40 source lines with the most warp-stall samples (share of samples / of executed instructions / of shared-memory
wavefronts, top stall reasons) and prints the same aggregated over the solver's phases (line ranges found from the function names)."""
import collections, csv, glob, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
cur, hdr, data = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 10 and r[0] == "Line No": hdr = r; continue
    if len(r) > 10 and hdr and r[0]:
        try: data.append((cur, int(r[0]), r))
        except ValueError: pass
ix = {h: i for i, h in enumerate(hdr)}
def g(r, h):
    try: return float(r[ix[h]])
    except (ValueError, KeyError): return 0.0
STALLS = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(g(r, "# Samples") for _, _, r in data) or 1.0
toti = sum(g(r, "Instructions Executed") for _, _, r in data) or 1.0
totw = sum(g(r, "L1 Wavefronts Shared") for _, _, r in data) or 1.0
def tops(r, s):
    st = sorted(((g(r, k), k[6:]) for k in STALLS), reverse=True)[:3]
    return " ".join("%s=%d%%" % (k, 100 * v / max(s, 1)) for v, k in st)
# phases = member functions of the solver, by the line they start at
src = open(glob.glob(os.path.join(ROOT, "online-non-linear-*", "csrc", "cmpc_solver.h"))[0]).read().split("\n")
starts = [(i, m.group(1)) for i, l in enumerate(src, 1) for m in [re.match(r"\s+CMPC_HD(?:_NOINLINE)? (?:static )?[\w:<>*& ]+? (\w+)\(", l)] if m]
marks = {i: n for i, l in enumerate(src, 1) for n, pat in (("pba", "M += [B A]' W (lower triangle)"), ("factor", "---- partial LDL' of the [u ; w] block"), ("gains/store", "---- gains: K = -L^-T")) if pat in l}
def phase(fn, ln):
    if fn != "cmpc_solver.h": return fn
    name = "?"
    for s, n in starts:
        if s <= ln: name = n
    if name == "backward":
        for s in sorted(marks):
            if s <= ln: name = marks[s]
    return name
ph = collections.defaultdict(lambda: collections.Counter())
for fn, ln, r in data:
    p = phase(fn, ln)
    ph[p]["s"] += g(r, "# Samples"); ph[p]["i"] += g(r, "Instructions Executed"); ph[p]["w"] += g(r, "L1 Wavefronts Shared")
    for k in STALLS: ph[p][k] += g(r, k)
print("%-22s %7s %7s %7s | top stalls (share of all samples)" % ("phase", "samp%", "inst%", "smemwf%"))
for p, d in sorted(ph.items(), key=lambda x: -x[1]["s"])[:18]:
    st = sorted(((d[k], k[6:]) for k in STALLS), reverse=True)[:4]
    print("%-22s %7.2f %7.2f %7.2f | %s" % (p, 100 * d["s"] / tot, 100 * d["i"] / toti, 100 * d["w"] / totw, " ".join("%s=%.1f" % (k, 100 * v / tot) for v, k in st)))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["file", "line", "samples_pct", "instructions_pct", "smem_wavefronts_pct", "top_stalls", "source"])
        for s, fn, ln, r in sorted(((g(r, "# Samples"), fn, ln, r) for fn, ln, r in data), reverse=True)[:40]:
            w.writerow([fn, ln, "%.2f" % (100 * s / tot), "%.2f" % (100 * g(r, "Instructions Executed") / toti), "%.2f" % (100 * g(r, "L1 Wavefronts Shared") / totw), tops(r, s), r[1].strip()[:160]])
