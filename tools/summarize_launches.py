"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the
LAST step (steps are delimited by the pair_concat kernel, launched once per forward)."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"(vsgg::[A-Za-z0-9_]+(<[^>]*>)?)", name)
    if m:
        return m.group(1)
    name = re.sub(r"\(.*", "", name)
    name = name.replace("at::native::", "")
    return name[:110]


def main(path, out=None):
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[start:]))
    marks = [i for i, r in enumerate(rows) if "pair_concat_kernel" in r["Kernel Name"]]
    # a step starts a few launches before pair_concat_fwd; use spacing between marks as the step length
    # window [marks[-2], marks[-1]) is exactly one step long (phase-shifted): same multiset of kernels as a step
    win = rows[marks[-2]:marks[-1]]
    agg = collections.OrderedDict()
    total = 0.0
    for r in win:
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    lines_out = ["one step = %d launches, %.3f ms summed kernel time (ncu-serialised, cold-cache)" % (len(win), total / 1e6)]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines_out.append("%8.3f ms  %5.1f%%  x%-4d %s" % (ns / 1e6, 100 * ns / total, n, k))
    txt = "\n".join(lines_out)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
