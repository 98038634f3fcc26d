"""Turn an .ncu-rep into a small text summary for profiles/ (run here, no GPU needed).

    python scripts/summarize_ncu.py gpurun_out/prof_gemm_tma.ncu-rep profiles/r01_gemm_tma_8192x8192x2048.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of {rep} (ncu --set full --clock-control none; per-launch values, cold-cache, serialised)"]
    for r in rows[2:]:
        lines.append("")
        for i, h in enumerate(hdr):
            if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                if r[i] not in ("", "0"):
                    lines.append(f"{h} [{units[i]}] = {r[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        si, ni, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        data = [(int(r[ni] or 0), r[si].strip(), r[ei]) for r in rows[2:] if len(r) > ni and (r[ni] or "0").isdigit()]
        tot = sum(d[0] for d in data) or 1
        lines.append("")
        lines.append(f"## top stall-sample SASS instructions of the first captured launch ({tot} samples, {len(data)} instructions)")
        for d in sorted(data, reverse=True)[:15]:
            lines.append(f"{100 * d[0] / tot:5.1f}%  exec={d[2]:>8}  {d[1][:100]}")
        ops = {}
        for _, s, _e in data:
            op = s.replace("@", " ").split()
            op = [t for t in op if not t.startswith("P") and not t.startswith("!")]
            if op:
                name = op[0].split(".")[0]
                ops[name] = ops.get(name, 0) + 1
        lines.append("")
        lines.append("## SASS opcode census (static): " + ", ".join(f"{k}={v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])
                                                                    if k in ("DMMA", "UTMALDG", "SYNCS", "DFMA", "LDS", "STS", "LDG", "STG", "CREDUX", "UTCHMMA", "LDGSTS", "BAR", "SHFL")))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, len(lines), "lines")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
