#!/usr/bin/env python
"""DRAM traffic of the Jacobian kernel from one `ncu --set full` capture (run here, no GPU needed):
   python tools/ncu_traffic.py gpurun_out/r02_jac.ncu-rep <observations> > profiles/r02_jacobian_traffic.json"""
import csv, io, json, subprocess, sys
rep, nobs = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
launches = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if "reproj_jac_fused_kernel" not in d.get("Kernel Name", ""):
        continue
    def val(k, want):
        v, u = float(d[k].replace(",", "")), units[hdr.index(k)]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}[u]
        return v * scale
    launches.append(dict(us=val("gpu__time_duration.sum", "us"), read=val("dram__bytes_read.sum", "byte"), write=val("dram__bytes_write.sum", "byte"),
                         dram_pct=float(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]), regs=int(d["launch__registers_per_thread"]),
                         warps_active_pct=float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]),
                         issue_active_pct=float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"])))
n = len(launches)
avg = lambda k: sum(l[k] for l in launches) / n
print(json.dumps(dict(kernel="reproj_jac_fused_kernel", capture=rep.split("/")[-1], launches=n, observations=nobs,
                      dram_bytes_read_per_launch=avg("read"), dram_bytes_written_per_launch=avg("write"), dram_bytes_per_launch=avg("read") + avg("write"),
                      bytes_per_observation=(avg("read") + avg("write")) / nobs, us_per_launch_under_ncu=avg("us"),
                      dram_gbs_under_ncu=(avg("read") + avg("write")) / avg("us") / 1e3, gpu_dram_throughput_pct=avg("dram_pct"),
                      registers=launches[0]["regs"], warps_active_pct=avg("warps_active_pct"), issue_active_pct=avg("issue_active_pct"),
                      note="ncu replays the kernel cold and serialised; bench.py's roofline times it live. The kernel writes 128 B per observation "
                           "(the -J_point translation block of the pose Jacobian is not materialised), reads the 32-byte record and gathers the point."), indent=1))
