"""gpurun_out/<tag>_ev_*.ncu-rep -> profiles/<tag>_kernel_evidence.md (one row per captured launch).
usage: python tools/evidence_table.py <tag> [report.ncu-rep ...]   (explicit reports replace the glob)"""
import csv, glob, io, json, os, subprocess, sys
tag = sys.argv[1]
# ALGORITHMIC bytes per launch of the HBM-bound kernels at the shapes tools/kernel_evidence.py runs:
# what the roofline fraction is quoted on; the DRAM counters next to it also depend on what the 126 MB L2 still holds.
_ROWS = 32 * 499 * 14
ALGO_BYTES = [
    ("add_ln_fwd_kernel<__nv_bfloat16", 3.0 * _ROWS * 256 * 2),          # x, residual in; y out
    ("add_ln_bwd_kernel<__nv_bfloat16", 4.0 * _ROWS * 256 * 2),          # dy, x, residual in; dres out (no dropout: dx = dres)
    ("add_ln_fwd_kernel<float", 3.0 * _ROWS * 256 * 4),
    ("add_ln_bwd_kernel<float", 4.0 * _ROWS * 256 * 4),
    ("frontend", 2048.0 * 499 * 40 * 4 + 2048.0 * 499 * 200 * 2),        # fp32 features in, 5-frame splice out in bf16
    ("cmvn_stats_kernel", 2048.0 * 450 * 40 * 4),                        # the real frames of every utterance, once
    ("adam_kernel", 7.0 * 35_000_000 * 4),                               # p, g, m, v in; p, m, v out
    ("attn_fwd_smallq_kernel", 250.0 * 499 * 256 * 4),                   # K and V of every utterance, once
]
def algo_bytes(name):
    for key, b in ALGO_BYTES:
        if key in name:
            return b
    return None
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
rows_out = []
for rep in (sys.argv[2:] or sorted(glob.glob("gpurun_out/%s_ev_*.ncu-rep" % tag))):
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
        ab = algo_bytes(d["Kernel Name"])
        rows_out.append((os.path.basename(rep), d["Kernel Name"].split("(")[0][:46], d["Grid Size"], dur_us, dram / 1e6,
                         dram / (dur_us * 1e-6) / 1e9 if dur_us else 0.0,
                         f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                         f("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                         f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                         int(f("launch__registers_per_thread")), bytes_of("l1tex__m_xbar2l1tex_read_bytes.sum") / 1e6,
                         (ab / (dur_us * 1e-6) / 1e9) if (ab and dur_us) else None))
with open("profiles/%s_kernel_evidence.md" % tag, "w") as fo:
    fo.write("# Per-kernel ncu evidence (`ncu --set full --clock-control none`, tools/gpu_evidence.sh, one launch per row)\n\n")
    fo.write("Shapes: tools/kernel_evidence.py.  Durations under ncu are cold-cache single launches; DRAM GB/s = (dram read + write bytes) / duration;\n")
    fo.write("measured peaks: HBM %.0f GB/s, bf16 %.0f TFLOP/s (MEASURED_PEAKS.json).\n\n" % (peaks["hbm_gbs"], peaks["bf16_tflops"]))
    fo.write("`algorithmic GB/s` = the kernel's algorithmic bytes at this shape (tools/evidence_table.py, DESIGN.md section 4) / duration: the\n")
    fo.write("roofline fraction of the HBM-bound kernels; the DRAM counters differ from it by what the 126 MB L2 holds at either end.\n\n")
    fo.write("| capture | kernel | grid | duration us | algorithmic GB/s | % of measured HBM | DRAM MB | DRAM GB/s | DRAM % of measured HBM | tensor pipe active % | SM throughput % | ncu DRAM % | regs | L2->SM MB |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows_out:
        alg = ("%.0f | %.0f%%" % (r[11], 100 * r[11] / peaks["hbm_gbs"])) if r[11] else "- | -"
        fo.write("| %s | %s | %s | %.1f | %s | %.1f | %.0f | %.0f%% | %.1f | %.1f | %.1f | %d | %.1f |\n" %
                 (r[0], r[1], r[2], r[3], alg, r[4], r[5], 100 * r[5] / peaks["hbm_gbs"], r[6], r[7], r[8], r[9], r[10]))
print(open("profiles/%s_kernel_evidence.md" % tag).read())
