"""Per-kernel SASS opcode histogram of the built library: the evidence that the kernels are Blackwell-native
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA 1-D), UTCBAR = tcgen05.commit, FFMA2 = fma.rn.f32x2,
SYNCS = mbarrier ops, DFMA = fp64 featurisation).  Runs on the CPU build box:
    python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import re, subprocess, sys
from collections import Counter, OrderedDict
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
lib = REPO / "neuralfoil_b200" / "lib" / "libneuralfoil_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FFMA", "HFMA2", "DFMA", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS", "ATOM", "RED", "BAR"]
kernels = OrderedDict()
cur = None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it, m.group(1))
        cur = re.sub(r"\((int|bool|unsigned int)\)", "", cur)
        cur = re.sub(r"\(nfb::.*\)$|\(float\*, int\)$", "", cur).replace("nfb::", "").replace("void ", "").replace(" >", ">")
        kernels[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["TOTAL"] += 1
        for k in KEYS:
            if op == k or (k in ("LDS", "STS", "LDG", "STG", "ATOM", "RED", "BAR", "MUFU", "SYNCS") and op.startswith(k)):
                kernels[cur][k] += 1
                break
arch = subprocess.run(["cuobjdump", "-lelf", str(lib)], capture_output=True, text=True).stdout
print(f"# {lib.name}: {sorted(set(re.findall(r'sm_[0-9a-z]+', arch)))}; cuobjdump -sass opcode counts per kernel (static instructions)")
print(f"# {'kernel':78s} " + " ".join(f"{k:>7s}" for k in ["TOTAL"] + KEYS))
tot = Counter()
for name, c in kernels.items():
    print(f"{name[:78]:78s}   " + " ".join(f"{c.get(k, 0):7d}" for k in ["TOTAL"] + KEYS))
    tot.update(c)
print(f"{'ALL KERNELS':80s} " + " ".join(f"{tot.get(k, 0):7d}" for k in ["TOTAL"] + KEYS))
