"""Aggregates `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_hot_lines.py src.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, cur_line, cur_src, hdr = None, None, "", None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None: continue
    if r[0] != "":
        cur_line, cur_src = r[0], r[1].strip()
    if len(r) <= 3 or r[2] == "": continue
    def g(name):
        try: return float(r[hdr[name]] or 0)
        except Exception: return 0.0
    key = (cur_file, cur_line)
    a = agg.setdefault(key, dict(src=cur_src, inst=0.0, smp=0.0, bar=0.0, lsb=0.0, ssb=0.0, wait=0.0, math=0.0, shw=0.0, shi=0.0))
    a["inst"] += g("Instructions Executed"); a["smp"] += g("# Samples"); a["bar"] += g("stall_barrier"); a["lsb"] += g("stall_long_sb")
    a["ssb"] += g("stall_short_sb"); a["wait"] += g("stall_wait"); a["math"] += g("stall_math")
    a["shw"] += g("L1 Wavefronts Shared"); a["shi"] += g("L1 Wavefronts Shared Ideal")
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["smp"] for a in agg.values())
print(f"total warp instructions {ti:.3e}, samples {ts:.0f}")
for a in agg.values(): a["busy"] = a["smp"] - a["bar"]
tb = sum(a["busy"] for a in agg.values())
print(f"non-barrier samples {tb:.0f} ({100*tb/ts:.1f}% of all)")
for title, key in (("samples", "smp"), ("non-barrier samples", "busy"), ("instructions", "inst")):
    print(f"--- top {topn} lines by {title}")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:topn]:
        print(f"{f}:{l:>4} busy {100*a['busy']/max(tb,1):5.1f}% smp {100*a['smp']/ts:5.1f}% inst {100*a['inst']/ti:5.1f}% [bar {100*a['bar']/max(a['smp'],1):3.0f}% lsb {100*a['lsb']/max(a['smp'],1):3.0f}% ssb {100*a['ssb']/max(a['smp'],1):3.0f}% wait {100*a['wait']/max(a['smp'],1):3.0f}%] shw/ideal {a['shw']/max(a['shi'],1):.1f} | {a['src'][:100]}")
