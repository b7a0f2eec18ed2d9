"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel summary CSV.
    python tools/launch_summary.py profiles/r01b_launch_list_bench_steps2.csv profiles/r01b_launch_list_summary.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        m = re.search(r"(\w+_kernel(?:<[^>(]*>)?)", r[kn])
        name = m.group(1) if m else re.sub(r"\(.*$", "", r[kn]).replace("void ", "")[:60]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{us:.1f},{us / tot:.4f}\n")
    print(open(dst).read())


if __name__ == "__main__":
    main(*sys.argv[1:3])
