"""Summarise `ncu -i X.ncu-rep --page raw --csv` (one row per profiled launch) into a small table."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"), ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("%-26s" % "kernel" + "".join("%11s" % n for _, n in WANT))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("nrb::", "")
        vals = []
        for m, n in WANT:
            v, u = r[idx[m]].replace(",", ""), units[idx[m]]
            try:
                f = float(v)
                if n == "time_us":
                    f = f * 1e3 if u == "ms" else (f / 1e3 if u == "ns" else f)
                if n.endswith("_MB"):
                    f = {"byte": f / 1e6, "Kbyte": f / 1e3, "Mbyte": f, "Gbyte": f * 1e3}.get(u, f)
                vals.append("%11.1f" % f)
            except ValueError:
                vals.append("%11s" % v[:10])
        print("%-26s" % name[:26] + "".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
