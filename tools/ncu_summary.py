#!/usr/bin/env python
"""Summarise an .ncu-rep (captured on the GPU box with `ncu --set full`) into the small text files kept under
profiles/:  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_r01   ->  x_r01.metrics.csv (one column
per captured launch, the metrics DESIGN.md quotes) and x_r01.hot_sass.txt (top stall/issue lines per launch,
when the report carries the source page).  Also:  --launches launches.csv out.md  for the per-launch time list."""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
    "smsp__average_warp_latency_issue_stalled_sleeping.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio",
    "smsp__average_warp_latency_issue_stalled_selected.ratio", "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
    "smsp__average_warp_latency_issue_stalled_imc_miss.ratio", "smsp__average_warp_latency_issue_stalled_drain.ratio",
    "smsp__average_warp_latency_issue_stalled_tex_throttle.ratio",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__ctas_launched.sum",
    "smsp__cycles_active.avg", "sm__cycles_active.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def metrics(rep, out_prefix):
    rows = ncu_csv(rep, "raw")
    H, units, data = rows[0], rows[1], rows[2:]
    ki = H.index("Kernel Name")
    with open(out_prefix + ".metrics.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        w.writerow(["kernel", ""] + [r[ki] for r in data])
        for m in METRICS:
            if m in H:
                i = H.index(m)
                w.writerow([m, units[i]] + [r[i] for r in data])
    print("wrote", out_prefix + ".metrics.csv")


def hot_sass(rep, out_prefix, top=30):
    rows = ncu_csv(rep, "source", ["--print-source", "sass"])
    blocks, hdr, name = [], None, ""
    for r in rows:
        if r and r[0] == "Kernel Name":
            name = r[1]
        elif r and r[0] == "Address":
            hdr = r
            blocks.append((name, hdr, []))
        elif hdr and len(r) == len(hdr):
            blocks[-1][2].append(r)
    # the page repeats each launch (one copy per source view); keep one
    blocks = [b for i, b in enumerate(blocks) if i == 0 or (b[0], b[2][:50]) != (blocks[i - 1][0], blocks[i - 1][2][:50])]
    if not blocks:
        return
    with open(out_prefix + ".hot_sass.txt", "w") as f:
        for n, (name, hdr, block) in enumerate(blocks):
            si, wi, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            tot = sum(float(r[wi] or 0) for r in block) or 1.0
            f.write(f"== launch {n}: {name[:110]}\n   {len(block)} SASS instructions, {int(tot)} warp-stall samples\n")
            st = sorted(((sum(float(r[i] or 0) for r in block), h) for i, h in stall_cols), reverse=True)
            f.write("   stall reasons (share of samples): " + ", ".join(f"{h[6:]} {100 * v / tot:.1f}%" for v, h in st if v / tot > 0.005) + "\n")
            mix = collections.Counter()
            for r in block:
                toks = r[si].split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                mix[op.split(".")[0]] += float(r[ii] or 0)
            f.write("   executed warp-instructions by opcode: " + ", ".join(f"{k} {int(v)}" for k, v in mix.most_common(22)) + "\n")
            f.write(f"   top {top} instructions by stall samples:\n")
            for r in sorted(block, key=lambda r: -float(r[wi] or 0))[:top]:
                f.write(f"   {100 * float(r[wi] or 0) / tot:6.2f}%  exec={r[ii]:>9}  {r[si].strip()}\n")
    print("wrote", out_prefix + ".hot_sass.txt")


def launches(csv_path, out_md):
    rows = [r for r in csv.reader(open(csv_path)) if r and (r[0] == "ID" or r[0].isdigit())]
    H = rows[0]
    ki, vi, gi, bi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki].split("(")[0][:100], [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out_md, "w") as f:
        f.write("| launches | avg us | share | grid | block | kernel |\n|---:|---:|---:|---|---|---|\n")
        for n, a in agg.items():
            f.write(f"| {a[0]} | {a[1] / a[0] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} | `{n}` |\n")
    print("wrote", out_md)


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        metrics(sys.argv[1], sys.argv[2])
        hot_sass(sys.argv[1], sys.argv[2])
