"""Static instruction mix of one kernel of a built library: python tools/sass_mix.py <lib.so> <kernel substring> [loop]
With "loop": only the instructions of the kernel's hottest loop body candidates are not identified -- the whole function is counted."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, mix, regs = None, collections.Counter(), 0
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur is None or pat not in cur:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", ln)
    if m:
        mix[(cur, m.group(1))] += 1
per = collections.defaultdict(collections.Counter)
for (f, op), n in mix.items():
    per[f][op] += n
for f, c in per.items():
    print(f, sum(c.values()), "instructions")
    for op, n in c.most_common(40):
        print(f"   {op:28s} {n}")
