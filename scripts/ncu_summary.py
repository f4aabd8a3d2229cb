"""Turn an ncu report (read here with `ncu -i ... --page raw --csv`) into the markdown tables kept under profiles/.

  python scripts/ncu_summary.py full  <raw.csv> <title>   -> per-launch metric table of a `--set full` capture
  python scripts/ncu_summary.py list  <launches.csv> <title> -> launch list of a `--metrics gpu__time_duration.sum` pass
"""
import csv
import sys
from collections import OrderedDict

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smsp__inst_executed_op_shared_atom.sum",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def full(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print("| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(vals))) + " | unit |")
    print("|---|" + "---|" * (len(vals) + 1))
    for m in WANT:
        if m in col:
            print(f"| {m} | " + " | ".join(v[col[m]] for v in vals) + f" | {units[col[m]]} |")
    rd, wr = col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum")
    if rd is not None:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = [float(v[rd]) * scale[units[rd]] + float(v[wr]) * scale[units[wr]] for v in vals]
        print("\nDRAM traffic per launch (read + write): " + ", ".join(f"{t / 1e6:.1f} MB" for t in tot))


def launches(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[col["Kernel Name"]]
        t = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        t_us = t / 1e3 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1e3)
        a = agg.setdefault(name, [0, 0.0, r[col["Grid Size"]], r[col["Block Size"]]])
        a[0] += 1
        a[1] += t_us
    total = sum(a[1] for a in agg.values())
    print(f"# {title}\n")
    print("| kernel | launches | grid | block | total us | mean us | share |")
    print("|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:100]}` | {a[0]} | {a[2]} | {a[3]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {100 * a[1] / total:.1f}% |")


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
