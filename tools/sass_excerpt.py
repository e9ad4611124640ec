"""profiles/r02_sass_excerpt.txt: per kernel of the built library, how often the design-relevant SASS instructions occur (TMA
loads, mbarrier waits, MATCH.ANY, MUFU.RSQ64H, 128-bit shared/global accesses, warp reductions ...) and their first occurrence.
python tools/sass_excerpt.py [lib.so] [out.txt]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "therldaisyworld_b200", "libdaisyworld_b200.so")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_sass_excerpt.txt")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTMALDG", "UTMASTG", "SYNCS", "MATCH.ANY", "MUFU.RSQ64H", "MUFU.RSQ ", "I2F.F64.U16", "I2F.U16", "REDUX", "LDS.128", "STS.128", "STG.E.128",
        "LDG.E.128", "SHFL", "VIMNMX", "DFMA", "FFMA", "BAR.SYNC", "BAR.ARV", "ATOMS", "ATOMG", "RED.E", "FENCE", "MEMBAR", "NANOSLEEP", "LDGSTS",
        "HMMA", "UTCMMA", "PRMT"]
cur, arch = None, None
per, first = collections.defaultdict(collections.Counter), collections.defaultdict(dict)
for ln in out.splitlines():
    m = re.search(r"arch = (\S+)", ln)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if not m:
        continue
    ins = m.group(2).strip() + " "
    for p in pats:
        if p in ins:
            per[cur][p.strip()] += 1
            first[cur].setdefault(p.strip(), f"/*{m.group(1)}*/ {ins.strip()}")
names = subprocess.run(["c++filt"] + list(per.keys()), capture_output=True, text=True).stdout.splitlines()
with open(dst, "w") as f:
    f.write(f"cuobjdump -sass {os.path.relpath(lib, ROOT)}  (arch {arch}); per kernel: count of the instructions that carry the Blackwell-specific /\n"
            "design-relevant paths, and the first occurrence of each (address + instruction). No tensor-core opcodes (HMMA / UTCMMA)\n"
            "anywhere: the path is a stencil. Regenerate: python tools/sass_excerpt.py\n\n")
    for k, d in zip(per.keys(), names):
        f.write(re.sub(r"\(.*", "", d) + "\n")
        for p, n in per[k].most_common():
            f.write(f"    {p:14s} x{n:<5d} {first[k][p]}\n")
        f.write("\n")
print(dst)
