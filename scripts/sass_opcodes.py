"""SASS opcode histogram per kernel of the built library -> profiles/r02_sass_opcodes.json (evidence that the hot kernels use
tcgen05 / TMEM / TMA: UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP =
cp.async.bulk, SYNCS = mbarrier).  Usage: python scripts/sass_opcodes.py [out.json]   (needs cuobjdump; no GPU)"""
import collections, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "audiocodecs_b200", "lib", "libaudiocodecs_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.json")
KEEP = ("FFMA", "MUFU", "STG", "LDG", "LD", "ST", "F2FP", "UTCHMMA", "LDTM", "STTM", "UTCBAR", "SYNCS", "STS", "LDS", "BAR", "ELECT",
        "UTCATOMSWS", "ATOMS", "ATOMG", "UTMALDG", "UTMASTG", "UBLKCP", "LDL", "STL")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        short = re.search(r"\d+([a-z_0-9]+_kernel)", name)
        cur = (short.group(1) if short else name)
        t = re.search(r"kernelI([A-Za-z0-9_]+?)E+v", name)   # template arguments, mangled
        if t:
            cur += "<" + t.group(1) + ">"
        kernels.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kernels[cur]["instructions"] += 1
        if m.group(1) in KEEP:
            kernels[cur][m.group(1)] += 1
totals = collections.Counter()
for c in kernels.values():
    for k, v in c.items():
        if k != "instructions":
            totals[k] += v
json.dump({"library": "audiocodecs_b200/lib/libaudiocodecs_b200.so (sm_100a)",
           "how": "python scripts/sass_opcodes.py: cuobjdump -sass, opcode histogram per kernel; LD / ST = generic loads / stores (zero in the "
                  "tap-GEMM kernels since the shared address space is kept: LDS / STS instead)",
           "totals": dict(totals), "kernels": {k: dict(v) for k, v in kernels.items()}}, open(out, "w"), indent=1)
print(out, dict(totals))
