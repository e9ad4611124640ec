"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name and share of the total."""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]); name = re.sub(r"<.*", "", name)
    t = float(r[14]); unit = r[13]
    t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] += t; cnt[name] += 1
total = sum(tot.values())
print(f"{len(rows)} launches, {total:.3f} ms of kernel time (cold-cache, serialised under ncu)")
for k, v in tot.most_common():
    print(f"  {k:40s} launches {cnt[k]:5d}  total {v:10.3f} ms  share {v / total:6.1%}  avg {v / cnt[k] * 1e3:10.1f} us")
