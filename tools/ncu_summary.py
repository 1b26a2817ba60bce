#!/usr/bin/env python3
"""One-screen summary of an ncu report (one kernel launch): duration, DRAM traffic and throughput against the measured HBM peak,
issue / pipe utilisation, shared-memory wavefronts, top stall reasons.   usage: ncu_summary.py <report.ncu-rep> [algorithmic_bytes]"""
import csv, json, os, subprocess, sys
rep = sys.argv[1]
alg = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6536.7
try:
    peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for vals in rows[2:]:
    m = {h: (v, u) for h, v, u in zip(hdr, vals, units)}

    def g(k, d=0.0):
        try:
            return float(m[k][0].replace(",", ""))
        except Exception:
            return d

    def scaled(k):  # value in base units (ncu prints Mbyte / Gbyte / msecond ...)
        v, u = g(k), m.get(k, ("", ""))[1].lower()
        for p, f in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0), ("msecond", 1e-3), ("usecond", 1e-6), ("nsecond", 1e-9), ("second", 1.0)):
            if u.startswith(p):
                return v * f
        return v * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    dur = scaled("gpu__time_duration.sum")
    rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
    print("kernel            :", m.get("Kernel Name", ("?",))[0][:100])
    print("grid x block      : %s x %s, %s regs/thread, %s KB dyn smem/block" % (m.get("Grid Size", ("?",))[0], m.get("Block Size", ("?",))[0],
          m.get("launch__registers_per_thread", ("?",))[0], m.get("launch__shared_mem_per_block_dynamic", ("?",))[0]))
    print("duration          : %.4f ms" % (dur * 1e3))
    print("DRAM read / write : %.1f MB / %.1f MB  -> %.1f GB/s = %.2f %% of the measured HBM peak (%.0f GB/s); ncu dram throughput %.1f %%" % (
        rd / 1e6, wr / 1e6, (rd + wr) / dur / 1e9, 100 * (rd + wr) / dur / 1e9 / peak, peak, g("dram__throughput.avg.pct_of_peak_sustained_elapsed", g("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed"))))
    if alg:
        print("algorithmic bytes : %.1f MB -> %.1f GB/s = %.3f of the HBM peak; DRAM traffic / algorithmic = %.2f" % (alg / 1e6, alg / dur / 1e9, alg / dur / 1e9 / peak, (rd + wr) / alg))
    print("issue active      : %.1f %%   warp-instructions %.3g   (ALU pipe %.1f %%, FMA %.1f %%, LSU %.1f %%, XU %.1f %%)" % (
        g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("smsp__inst_executed.sum"), g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")))
    print("shared wavefronts : %.3g (%.1f %% of the LSU data pipe), bank conflicts %.3g" % (g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
          g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")))
    print("occupancy         : %.1f %% of max warps" % g("sm__warps_active.avg.pct_of_peak_sustained_active"))
    st = sorted(((k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), g(k)) for k in m if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), key=lambda kv: -kv[1])
    print("stalls / issue    :", ", ".join("%s %.2f" % kv for kv in st[:6]))
    print()
