"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into a compact table: one row per profiled launch with the metrics
DESIGN.md / profiles/README.md quote.  Usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tc.sum", "tc_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__cycles_active.avg", "cycles")]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    for path in sys.argv[1:]:
        hdr, units, rows = rows_of(path)
        cols = [(hdr.index(m), short) for m, short in WANT if m in hdr]
        ki = hdr.index("Kernel Name")
        print("## %s" % path)
        print("| kernel | " + " | ".join("%s [%s]" % (s, units[i]) if units[i] else s for i, s in cols) + " |")
        print("|---|" + "---|" * len(cols))
        for r in rows:
            name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
            print("| %s | " % name + " | ".join(r[i] for i, _ in cols) + " |")
        print()


if __name__ == "__main__":
    main()
