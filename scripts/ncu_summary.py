"""summarise `ncu --page raw --csv` exports (scripts/ncu_capture.sh) into the handful of numbers DESIGN.md / bench.py quote:
    python scripts/ncu_summary.py gpurun_out/r02_lic_sample_cfg3.raw.csv [...]  ->  JSON on stdout"""
import csv
import json
import sys

KEYS = {
    "duration_ms": ("gpu__time_duration.sum", 1e-6),            # ns
    "warp_instructions": ("smsp__inst_executed.sum", 1),
    "issue_slots_busy_pct": ("sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
    "ipc_issued_per_sm": ("sm__inst_issued.avg.per_cycle_active", 1),
    "fma_pipe_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    "fmaheavy_pipe_pct": ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "alu_pipe_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    "xu_pipe_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    "lsu_pipe_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
    "l1_data_pipe_pct": ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
    "l1_sector_hit_pct": ("l1tex__t_sector_hit_rate.pct", 1),
    "l2_sector_hit_pct": ("lts__t_sector_hit_rate.pct", 1),
    "dram_read_bytes": ("dram__bytes_read.sum", 1),
    "dram_write_bytes": ("dram__bytes_write.sum", 1),
    "dram_pct_of_peak": ("dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "l2_throughput_pct": ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "l1_throughput_pct": ("l1tex__throughput.avg.pct_of_peak_sustained_active", 1),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "registers_per_thread": ("launch__registers_per_thread", 1),
    "sm_cycles_active_avg": ("sm__cycles_active.avg", 1),
    "stall_math_pipe_throttle": ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1),
    "stall_not_selected": ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", 1),
    "stall_long_scoreboard": ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1),
    "stall_short_scoreboard": ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1),
    "stall_wait": ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1),
    "stall_lg_throttle": ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", 1),
    "stall_dispatch": ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", 1),
    "stall_no_instruction": ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", 1),
    "stall_branch_resolving": ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", 1),
}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return v


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    out = {}
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name", "")[:90]}
        for k, (m, scale) in KEYS.items():
            if m in d and d[m] != "":
                v = num(d[m])
                if isinstance(v, float):
                    un = u.get(m, "")
                    if k == "duration_ms":
                        v = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(un, 1e-6)
                    elif k.endswith("_bytes"):
                        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(un, 1)
                    rec[k] = v
        out[path.split("/")[-1].replace(".raw.csv", "")] = rec
    return out


if __name__ == "__main__":
    res = {}
    for p in sys.argv[1:]:
        res.update(summarise(p))
    print(json.dumps(res, indent=1))
