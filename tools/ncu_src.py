"""Hot spots of one kernel in an .ncu-rep, per CUDA source line (run here, no GPU needed).
usage: python tools/ncu_src.py <rep> <kernel-regex> [top-N]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
agg = collections.OrderedDict()
fname, func, hdr = "", "", None
first_func = None
for i, l in enumerate(lines):
    if l.startswith('"File Path"'): fname = next(csv.reader([l]))[1].split("/")[-1]; hdr = None; continue
    if l.startswith('"Function Name"'):
        func = next(csv.reader([l]))[1]
        if first_func is None: first_func = func
        continue
    if l.startswith('"Line No"'): hdr = next(csv.reader([l])); continue
    if hdr is None or func != first_func: continue
    r = next(csv.reader([l]))
    if len(r) != len(hdr): continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    if d.get("Address", "-") != "-": continue      # per-line rows only (aggregated over their SASS)
    key = (fname, d["Line No"])
    e = agg.setdefault(key, dict(src=d["Source"], samples=0.0, inst=0.0, st=collections.Counter()))
    e["samples"] += float(d["# Samples"] or 0); e["inst"] += float(d["Instructions Executed"] or 0)
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k: e["st"][k[6:]] += float(d[k] or 0)
tot = sum(e["samples"] for e in agg.values())
allst = collections.Counter()
for e in agg.values(): allst.update(e["st"])
print(first_func[:150])
print("samples", tot, "inst", sum(e["inst"] for e in agg.values()), {k: round(100 * v / max(tot, 1), 1) for k, v in allst.most_common(8)})
for (f, ln), e in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    why = " ".join(f"{k}:{100*v/max(e['samples'],1):.0f}" for k, v in e["st"].most_common(2))
    print(f"{100*e['samples']/tot:5.1f}% {f}:{ln:>5s} inst={e['inst']:9.0f} {e['src'].strip()[:100]:100s} {why}")
