"""Blackwell-native evidence: per-kernel counts of the tcgen05 / TMEM / TMA SASS opcodes in the built library
(cuobjdump -sass; run here, no GPU needed):
    python tools/sass_histogram.py > profiles/sass_rNN_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vit-2spn_b200", "libvit2spn.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "HMMA", "MUFU", "FENCE.VIEW.ASYNC"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("void ", "").replace("v2s::(anonymous namespace)::", "").replace("v2s::", "")
        name = re.sub(r"\((?!anonymous).*", "", name)
        kern = name
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        op = m.group(1)
        counts[kern]["_all"] += 1
        for o in OPS:
            if op.startswith(o):
                counts[kern][o] += 1
                total[o] += 1
print(f"# {LIB}: cubin architectures {arch}; SASS opcode counts per kernel (tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR,")
print("# tcgen05.ld / st = LDTM / STTM, TMA = UTMALDG / UTMASTG / UTMAREDG, mbarrier = SYNCS; HMMA would be the legacy mma.sync path)")
print("# totals: " + "  ".join(f"{o}={total[o]}" for o in OPS))
print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{o[:8]:>8s}" for o in OPS))
for k, c in counts.items():
    if any(c[o] for o in OPS[:8]):
        print(f"{k[:78]:78s} {c['_all']:6d} " + " ".join(f"{c[o]:8d}" for o in OPS))
