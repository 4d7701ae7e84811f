"""Per-kernel counts of the Blackwell tensor-core / TMA / TMEM SASS opcodes in libpka_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), UTMALDG / UTMASTG (TMA bulk tensor load / store), LDTM / STTM (tcgen05.ld /
tcgen05.st), UTCBAR (tcgen05.commit), MUFU.EX2.  usage: python tools/sass_opcodes.py > profiles/sass_opcodes_r02.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pytorch-kaldi-asr_b200", "csrc", "libpka_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.*2CTA", "UTMALDG.*MULTICAST", "UTMASTG", "LDTM", "STTM", "UTCBAR", "MUFU.EX2", "SYNCS"]
rows, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        rows[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for op in ops:
        if re.search(r"\b" + op.replace(".", r"\.").replace(r"\.*", ".*") + r"\b", line) or (op.endswith("2CTA") and re.search(op.replace(".", r"\.").replace(r"\.*", ".*"), line)):
            rows[cur][op] += 1
print("SASS opcode counts per kernel, libpka_b200.so (sm_100a), kernels with at least one tensor-core / TMA / TMEM opcode")
print("%-58s" % "kernel" + "".join("%12s" % o.replace(".*", "..") for o in ops))
tot = collections.Counter()
for k, c in rows.items():
    if any(c[o] for o in ops[:9]):
        print("%-58s" % k[:58] + "".join("%12d" % c[o] for o in ops))
        tot.update(c)
print("%-58s" % "total" + "".join("%12d" % tot[o] for o in ops))
