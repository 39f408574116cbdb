"""Turns an .ncu-rep (read here, no GPU needed) into the text summary committed under profiles/.
    python tools/summarize_ncu.py gpurun_out/prof_chain_r1.ncu-rep profiles/r01_chain_kernel_ncu.md "title"
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def ncu(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(io.StringIO(ncu([rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    L = [f"# {title}", "", f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, one launch).",
         f"Kernel: `{m.get('Kernel Name', ('?', ''))[0]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            L.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    L += ["", "## Warp stall reasons (per issue-active cycle)", "", "| reason | ratio |", "|---|---|"]
    st = [(float(v[0]), h) for h, v in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:10]:
        L.append(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |")

    src = list(csv.reader(io.StringIO(ncu([rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    hdr2 = next((r for r in src if len(r) > 3 and r[0] == "Line No"), None)
    if hdr2:
        ix = {h: i for i, h in enumerate(hdr2)}
        cur, lines, tot_i, tot_s = None, [], 0, 0
        for r in src:
            if len(r) >= 2 and r[0] == "File Path":
                cur = r[1].split("/")[-1]
                continue
            if len(r) < len(hdr2) or r[0] in ("Line No", "") or r[2] != "-":
                continue
            try:
                n, s, b = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), int(r[ix["stall_barrier"]])
            except ValueError:
                continue
            lines.append((n, s, b, cur, r[0], r[1].strip()[:100]))
            tot_i += n; tot_s += s
        L += ["", "## Hottest source lines (warp-level instructions executed; warp samples; of which at a barrier)", "",
              "| % inst | % samples | % samples at barrier | line | source |", "|---|---|---|---|---|"]
        for n, s, b, f, ln, text in sorted(lines, reverse=True)[:25]:
            L.append(f"| {100 * n / tot_i:.2f} | {100 * s / tot_s:.2f} | {100 * b / tot_s:.2f} | {f}:{ln} | `{text.replace('|', '/')}` |")
        ops = {}
        for r in src:
            if len(r) >= len(hdr2) and r[0] == "" and r[2] not in ("-", "Address"):
                toks = r[3].split()
                op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
                try:
                    ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[ix["Instructions Executed"]])
                except ValueError:
                    pass
        tot = sum(ops.values()) or 1
        L += ["", "## SASS opcode mix (warp-level instructions executed)", "", "| opcode | % |", "|---|---|"]
        for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:16]:
            L.append(f"| {op} | {100 * n / tot:.2f} |")
    open(out, "w").write("\n".join(L) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
