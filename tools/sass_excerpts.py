#!/usr/bin/env python
"""cuobjdump -sass of libobvi_ba.so, reduced to the evidence lines: per kernel the count of the mnemonics that prove the
Blackwell-side mechanisms (UBLKCP = cp.async.bulk / TMA 1-D copies, SYNCS = mbarrier, DMMA = fp64 tensor cores, REDG =
fire-and-forget reductions, LDGSTS = cp.async, WARPSYNC / SHFL / REDUX = warp collectives) plus the first occurrence of each."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "obvi-slam_b200/libobvi_ba.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "DMMA", "REDG", "ATOMG", "LDGSTS", "REDUX", "SHFL", "BAR.SYNC", "LDS", "STS", "LDG", "STG", "DFMA", "UTC", "LDTM"]
fn, stats, first = None, collections.OrderedDict(), {}
arch = set(re.findall(r"arch = (sm_\w+)", out))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        stats[fn] = collections.Counter(); continue
    if fn is None: continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if not m: continue
    op = m.group(1); stats[fn]["_total"] += 1
    for k in KEYS:
        if op.startswith(k):
            stats[fn][k] += 1
            first.setdefault((fn, k), re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", line).strip())
print("library:", so, " architectures in the fatbinary:", sorted(arch))
for fn, c in stats.items():
    if c["_total"] < 40: continue
    print(f"\n== {fn}   ({c['_total']} SASS instructions)")
    print("   " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
    for k in ("UBLKCP", "SYNCS", "DMMA", "REDG", "LDGSTS", "REDUX"):
        if (fn, k) in first: print(f"   first {k:7s}: {first[(fn, k)][:140]}")
