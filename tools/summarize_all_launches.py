"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list over ALL captured launches (per-kernel totals)."""
import collections, csv, re, sys
def short(name):
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"((vsgg|g2)::[A-Za-z0-9_:]+(<[^>]*>)?)", name)
    if m:
        return m.group(1)
    return re.sub(r"\(.*", "", name).replace("at::native::", "")[:100]
lines = open(sys.argv[1]).readlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[start:]))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
agg, total = collections.OrderedDict(), 0.0
for r in rows:
    ns = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("us", "usecond"): ns *= 1e3
    elif r["Metric Unit"] in ("ms", "msecond"): ns *= 1e6
    a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0]); a[0] += 1; a[1] += ns; total += ns
print("%d launches, %.3f ms summed kernel time (ncu-serialised, cold-cache), divided by %g" % (len(rows), total / 1e6 / div, div))
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%9.3f ms %5.1f%%  x%-4d %s" % (ns / 1e6 / div, 100 * ns / total, c / div, k))
