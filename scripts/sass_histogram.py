"""SASS opcode histogram of the in-tree library (evidence that the hot kernels are Blackwell-native,
B200_PROFILING.md "What proves a Blackwell-native kernel"): per kernel, the counts of the tensor-core / TMEM / TMA /
async-copy / atomic mnemonics.  usage: python scripts/sass_histogram.py [lib.so] > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parents[1] / "nerf_downstream_b200" / "libsparseconv_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMACMDFLUSH",
         "LDGSTS", "SYNCS", "ELECT", "REDG", "RED", "ATOMG", "ATOM", "HMMA", "LDG", "STG", "LDS", "STS", "SHFL", "VOTE", "BAR"]
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if cur and m:
        funcs[cur][m.group(1)] += 1
        funcs[cur]["_total"] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib}")
print("# arch: " + ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))))
for (name, c), dem in zip(funcs.items(), demangled):
    short = re.sub(r"\(.*", "", dem).replace("void ", "")
    hits = [(k, c[k]) for k in WATCH if c.get(k)]
    print(f"{short:70s} {c['_total']:6d} instr  " + "  ".join(f"{k}={v}" for k, v in hits))
