#!/usr/bin/env python
"""Writes a short, committed summary (markdown) of an .ncu-rep: the counters the roofline and DESIGN.md quote.

usage: ncu_summary.py report.ncu-rep [out.md] [--alg-bytes N]
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed).
"""
import csv
import io
import subprocess
import sys

CURATED = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rep = args[0]
    out = args[1] if len(args) > 1 else None
    alg = None
    if "--alg-bytes" in sys.argv:
        alg = float(sys.argv[sys.argv.index("--alg-bytes") + 1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu summary of `%s`" % rep.split("/")[-1], "",
             "Captured with `ncu --set full --clock-control none --import-source on` on a B200 (gpurun); numbers under a",
             "profiler are cold-cache/serialised and are NOT bench values.", ""]
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        lines.append("## %s  (launch id %s)" % (d.get("Kernel Name", ("?",))[0], d.get("ID", ("?",))[0]))
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for m in CURATED:
            if m in d and d[m][0] != "":
                lines.append("| %s | %s | %s |" % (m, d[m][0], d[m][1]))
        try:
            t = float(d["gpu__time_duration.sum"][0])
            tu = d["gpu__time_duration.sum"][1]
            t_s = t * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(tu, 1e-3)
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            rd = float(d["dram__bytes_read.sum"][0]) * scale[d["dram__bytes_read.sum"][1]]
            wr = float(d["dram__bytes_write.sum"][0]) * scale[d["dram__bytes_write.sum"][1]]
            lines.append("| derived: DRAM traffic (read+write) | %.4g | byte |" % (rd + wr))
            lines.append("| derived: DRAM GB/s under ncu | %.1f | GB/s |" % ((rd + wr) / t_s / 1e9))
            if alg:
                lines.append("| derived: traffic / algorithmic bytes (%.4g) | %.2f | x |" % (alg, (rd + wr) / alg))
        except Exception:
            pass
        stalls = []
        for h in hdr:
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") or (
                    "smsp__average_warp_latency_issue_stalled" in h and h.endswith(".ratio")):
                try:
                    stalls.append((float(d[h][0]), h))
                except Exception:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            lines.append("")
            lines.append("Top warp-stall reasons (cycles per issued instruction):")
            lines.append("")
            for v, h in stalls[:8]:
                lines.append("- %s = %.2f" % (h.replace("smsp__average_", ""), v))
        lines.append("")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
