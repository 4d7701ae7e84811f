"""gpurun_out/<tag>_ev_*.ncu-rep -> profiles/<tag>_kernel_evidence.md (one row per captured launch)."""
import csv, glob, io, json, os, subprocess, sys
tag = sys.argv[1]
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
rows_out = []
for rep in sorted(glob.glob("gpurun_out/%s_ev_*.ncu-rep" % tag)):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        def f(k, default=0.0):
            try:
                return float(d.get(k, default))
            except ValueError:
                return default
        def bytes_of(k):
            v, un = f(k), u.get(k, "")
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(un, 1)
        dur_us = f("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u.get("gpu__time_duration.sum", "us"), 1)
        dram = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
        rows_out.append((os.path.basename(rep), d["Kernel Name"].split("(")[0][:46], d["Grid Size"], dur_us, dram / 1e6,
                         dram / (dur_us * 1e-6) / 1e9 if dur_us else 0.0,
                         f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                         f("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                         f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                         int(f("launch__registers_per_thread")), bytes_of("l1tex__m_xbar2l1tex_read_bytes.sum") / 1e6))
with open("profiles/%s_kernel_evidence.md" % tag, "w") as fo:
    fo.write("# Per-kernel ncu evidence (`ncu --set full --clock-control none`, tools/gpu_evidence.sh, one launch per row)\n\n")
    fo.write("Shapes: tools/kernel_evidence.py.  Durations under ncu are cold-cache single launches; DRAM GB/s = (dram read + write bytes) / duration;\n")
    fo.write("measured peaks: HBM %.0f GB/s, bf16 %.0f TFLOP/s (MEASURED_PEAKS.json).\n\n" % (peaks["hbm_gbs"], peaks["bf16_tflops"]))
    fo.write("| capture | kernel | grid | duration us | DRAM MB | DRAM GB/s | % of measured HBM | tensor pipe active % | SM throughput % | ncu DRAM % | regs | L2->SM MB |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows_out:
        fo.write("| %s | %s | %s | %.1f | %.1f | %.0f | %.0f%% | %.1f | %.1f | %.1f | %d | %.1f |\n" %
                 (r[0], r[1], r[2], r[3], r[4], r[5], 100 * r[5] / peaks["hbm_gbs"], r[6], r[7], r[8], r[9], r[10]))
print(open("profiles/%s_kernel_evidence.md" % tag).read())
