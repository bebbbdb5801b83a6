"""Developer tool: turn the outputs of `scripts/gpu_r2.sh <tag>` (gpurun_out/<tag>_*) into the tracked summaries under profiles/.

usage: python scripts/make_profiles.py <tag> [prefix]        (prefix defaults to r2)
writes  profiles/<prefix>_bench_n1.json, _bench_ref.json, _launches.csv/.txt, _ncu_c3.txt, _ncu_c4.txt, _parity.txt and the N = 1
entries of profiles/traffic.json.  Needs `ncu` (to export the .ncu-rep files as CSV); no GPU.
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

TABLE = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
         ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
         ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
         ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
         ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
         ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"), ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
         ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst")]
EXTRA = [("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-stage wavefronts % of peak"),
         ("l1tex__data_pipe_lsu_wavefronts.sum", "L1 data-stage wavefronts"),
         ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "L1 tag-stage wavefronts, global loads"),
         ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
         ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
         ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
         ("smsp__inst_executed.sum", "warp instructions"),
         ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local-memory load sectors"),
         ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "local-memory store sectors"),
         ("lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "L2 read-miss sectors from L1/TEX (x32 B = DRAM reads they cause)"),
         ("lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum", "L2 write-miss sectors from L1/TEX"),
         ("lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum", "L2 RED-miss sectors from L1/TEX"),
         ("lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum", "L2 miss sectors arriving over the die-to-die fabric")]


def short(name):
    name = name.replace("void ", "").replace("nrb::", "")
    return name.split("(")[0]


def ncu_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, [r for r in rows[2:] if len(r) >= len(hdr)]


def ncu_summary(tag, cfg, header, dst):
    rep = os.path.join(OUT, "%s_prof_%s.ncu-rep" % (tag, cfg))
    if not os.path.exists(rep):
        print("skip", rep)
        return None
    hdr, units, rows = ncu_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    kn = idx["Kernel Name"]

    def val(r, key):
        if key not in idx:
            return float("nan")
        v = float(r[idx[key]].replace(",", "") or "nan")
        u = units[idx[key]]
        if key == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        if key.startswith("dram__bytes"):
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        return v

    lines = list(header)
    lines.append("%-28s" % "kernel" + "".join("%11s" % n for _, n in TABLE))
    for r in rows:
        lines.append("%-28s" % short(r[kn])[:28] + "".join("%11.1f" % val(r, k) for k, _ in TABLE))
    lines.append("")
    lines.append("%-76s" % "metric" + "".join("%24s" % short(r[kn])[:22] for r in rows))
    for k, label in EXTRA:
        if k in idx:
            lines.append("%-76s" % label + "".join("%24.1f" % val(r, k) for r in rows))
    open(dst, "w").write("\n".join(lines) + "\n")
    traffic = [(short(r[kn]), (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) * 1e6) for r in rows]
    return traffic


def launches(tag, prefix):
    src = os.path.join(OUT, "%s_launches.csv" % tag)
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, "%s_launches.csv" % prefix))
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row.get("Metric Unit", "ns"), 1e-3)
        k = short(row["Kernel Name"])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + v)
    total = sum(t for _, t in agg.values())
    out = ["# %s — ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extra none`; per-launch times are "
           "cold-cache and serialised (ncu): compare SHARES" % os.path.basename(src),
           "%-48s %6s %12s %7s" % ("kernel", "n", "total_us", "share")]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-48s %6d %12.1f %7.3f" % (k[:48], n, t, t / total))
    open(os.path.join(PROF, "%s_launches.txt" % prefix), "w").write("\n".join(out) + "\n")


def parity(tag, prefix):
    src = os.path.join(OUT, "%s_parity.jsonl" % tag)
    if not os.path.exists(src):
        return
    passed = ""
    log = os.path.join(OUT, "%s_pytest.log" % tag)
    if os.path.exists(log):
        passed = [l for l in open(log) if " passed" in l][-1].strip()
    out = ["# parity report of `pytest tests -m gpu` (%s), B200, end of round 2 — device (through the C-ABI) vs the f64 oracle / the "
           "brute-force checker" % passed,
           "# tolerance: per-channel |delta| <= 1/255; bar: <= 1e-3 of the pixels over it; every over-tolerance pixel classified "
           "(tests/util.py): edge flip / precision-sensitive (the oracle's own f32 mode moves there too) / unexplained",
           "%-58s %9s %6s %10s %6s %6s %11s %8s %9s" % ("frame", "pixels", "over", "frac_over", "flips", "sens.", "unexplained", "max_abs", "mean_abs")]
    for l in open(src):
        d = json.loads(l)
        out.append("%-58s %9d %6d %10.2e %6d %6s %11d %8.3f %9.2e" % (
            d["what"][:58], d["pixels"], d["over"], d["frac_over"], d["edge_flips"],
            d.get("precision_sensitive", "-"), d["unexplained"], d["max_abs"], d["mean_abs"]))
    open(os.path.join(PROF, "%s_parity.txt" % prefix), "w").write("\n".join(out) + "\n")


def main():
    tag = sys.argv[1]
    prefix = sys.argv[2] if len(sys.argv) > 2 else "r2"
    for a, b in (("bench.json", "bench_n1.json"), ("bench_ref.json", "bench_ref.json")):
        src = os.path.join(OUT, "%s_%s" % (tag, a))
        if os.path.exists(src):
            shutil.copy(src, os.path.join(PROF, "%s_%s" % (prefix, b)))
    launches(tag, prefix)
    parity(tag, prefix)
    tj_path = os.path.join(PROF, "traffic.json")
    tj = json.load(open(tj_path))
    hdr3 = ["# ncu --set full of ONE C3 frame (crytek_sponza stand-in 1920x1080x4spp): the seven launches from the shade of the previous "
            "frame's wave 0 on, in launch order (scripts/gpu_r2.sh, capture %s)" % tag,
            "# per-launch times are cold-cache and serialised under ncu: compare shares, not absolutes"]
    t3 = ncu_summary(tag, "c3", hdr3, os.path.join(PROF, "%s_ncu_c3.txt" % prefix))
    hdr4 = ["# ncu --set full of C4 (hairball stand-in, 2.88 M triangles, 1920x1080x8spp, one batch): launches 3-5 of the second frame "
            "(scripts/gpu_r2.sh, capture %s)" % tag,
            "# per-launch times are cold-cache and serialised under ncu: compare shares, not absolutes"]
    t4 = ncu_summary(tag, "c4", hdr4, os.path.join(PROF, "%s_ncu_c4.txt" % prefix))
    for cfg, tr in (("C3", t3), ("C4", t4)):
        if not tr:
            continue
        v = [b for k, b in tr if k.startswith("trace_kernel")]
        if v:
            tj[cfg]["1"]["trace_kernel"] = sum(v) / len(v)
            tj[cfg]["1"]["launches_averaged"] = len(v)
            tj[cfg]["1"]["source"] = "profiles/%s_ncu_%s.txt" % (prefix, cfg.lower())
    json.dump(tj, open(tj_path, "w"), indent=1)
    print("profiles updated from", tag)


if __name__ == "__main__":
    main()
