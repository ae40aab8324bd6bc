"""ncu report -> the handful of raw metrics quoted in DESIGN.md (one 'name [unit] = value' line each).

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep 'kernel name regex' > profiles/rNN_<kernel>_ncu_summary.txt
"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|launch__(grid_size|block_size|registers_per_thread|"
                  r"shared_mem_per_block_dynamic)|sm__cycles_active\.avg|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__pipe_(fma|alu)_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__inst_executed_pipe_(xu|lsu)\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__icc_request_hit_rate\.pct|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed)$")


def main(rep, pattern):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if not re.search(pattern, r[ki]):
            continue
        print("Kernel Name [] = %s" % r[ki])
        for h, u, v in zip(hdr, units, r):
            if KEEP.match(h):
                print("%s [%s] = %s" % (h, u, v))
        print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ".")
