"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/<tag>_launches.csv profiles/<tag>_launches.txt
    python tools/summarize_ncu.py full gpurun_out/<tag>_full_<kernel>.ncu-rep profiles/<tag>_ncu_<kernel>.txt
"""
import collections, csv, io, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "launch__waves_per_multiprocessor", "launch__cluster", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"]


def launches(src, dst):
    lines = open(src).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[i:]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = re.sub(r"\(.*", "", r["Kernel Name"])[:90]
        agg[(n, r["Grid Size"], r["Block Size"])][0] += 1
        agg[(n, r["Grid Size"], r["Block Size"])][1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES)\n")
        f.write("# source: %s, %d launches, %.1f us total\n" % (src, len(rows), tot / 1e3))
        f.write("# %10s %6s %9s %6s  kernel grid block\n" % ("total_us", "count", "avg_us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%12.1f %6d %9.2f %5.1f%%  %s %s %s\n" % (v[1] / 1e3, v[0], v[1] / v[0] / 1e3, 100 * v[1] / tot, k[0], k[1], k[2]))


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; source: %s\n" % src)
        for r in rows[2:]:
            rec = dict(zip(hdr, r))
            f.write("\n== %s  grid %s block %s\n" % (rec.get("Kernel Name", "?")[:100], rec.get("Grid Size"), rec.get("Block Size")))
            for h, u, val in zip(hdr, units, r):
                if any(h.startswith(k) for k in KEYS):
                    f.write("  %-80s %s %s\n" % (h, val, u))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
