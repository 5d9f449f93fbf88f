"""Summaries of ncu outputs for profiles/.

  ncu_summary.py launches <raw launch-list csv> <out csv> [note]
      per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list
  ncu_summary.py full <out csv> <report.ncu-rep> [<report.ncu-rep> ...]
      selected metrics of every launch in `ncu --set full` reports (read with `ncu -i ... --page raw --csv`)"""
import csv, io, subprocess, sys
from collections import OrderedDict

PLUMBING = "torch / CUB plumbing (sorts, scans, index building)"


def launches(raw, out, note=""):
    rows = [r for r in csv.reader(l for l in open(raw) if l.startswith('"'))]
    h = rows[0]
    iK, iV = h.index("Kernel Name"), h.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[iK].split("(")[0]
        if not name.startswith("xmap::") and not name.startswith("xsim_"):
            name = PLUMBING
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += float(r[iV].replace(",", "")) / 1e6
    tot = sum(a[1] for a in agg.values()); n = sum(a[0] for a in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share_of_captured_time_pct"])
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, c, "%.3f" % ms, "%.1f" % (100 * ms / tot)])
        w.writerow(["# " + note, n, "%.3f" % tot, 100])


FULL = OrderedDict([
    ("time_ms", "gpu__time_duration.sum"),
    ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("ipc", "sm__inst_executed.avg.per_cycle_active"),
    ("regs", "launch__registers_per_thread"), ("dyn_smem_kb", "launch__shared_mem_per_block_dynamic"),
    ("lsu_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("smem_atom_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum"),
    ("smem_atom_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("thread_inst_per_warp_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
])


def full(out, reports):
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block"] + list(FULL) + ["units as printed by ncu (value unit)"])
        for rep in reports:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(txt)))
            h, units = rows[0], rows[1]
            col = {n: i for i, n in enumerate(h)}
            for r in rows[2:]:
                line = [r[col["Kernel Name"]].split("(")[0], r[col["Grid Size"]], r[col["Block Size"]]]
                for k, m in FULL.items():
                    i = col.get(m)
                    line.append("" if i is None else ("%s %s" % (r[i], units[i])).strip())
                w.writerow(line)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3:])
