"""Turn an .ncu-rep (ncu --set full) into a small text summary for profiles/: key metrics per captured kernel."""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    r"^gpu__time_duration\.sum$", r"^launch__grid_size$", r"^launch__block_size$", r"^launch__registers_per_thread$",
    r"^launch__occupancy_limit_(registers|shared_mem|warps|blocks)$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__inst_executed\.sum$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^sm__inst_executed\.avg\.per_cycle_elapsed$",
    r"^smsp__thread_inst_executed_per_inst_executed\.ratio$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__inst_executed_pipe_(fma|fmaheavy|fmalite|alu|xu|lsu|fp64|uniform)\.sum$", r"^sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_fma\.avg\.pct_of_peak_sustained_active$", r"^sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
    r"^smsp__warps_eligible\.avg\.per_cycle_active$", r"^smsp__warps_active\.avg\.per_cycle_active$",
    r"^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$", r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^lts__t_bytes\.sum$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
    r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$", r"^smsp__sass_average_branch_targets_threads_uniform\.pct$",
    r"^sm__cycles_elapsed\.max$", r"^smsp__cycles_active\.avg$",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    pats = [re.compile(k) for k in KEYS]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        print("## kernel: %s   grid %s block %s" % (d.get("Kernel Name", "?"), d.get("launch__grid_size"), d.get("launch__block_size")))
        for h, u, v in zip(hdr, units, row):
            if any(p.search(h) for p in pats):
                print("%-82s %18s %s" % (h, v, u))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
