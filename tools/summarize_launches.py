"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total time, share."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    body = rows[rows.index(hdr) + 1:]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in body:
        if len(r) <= v:
            continue
        val = float(r[v].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(r[u], 1e-3)
        name = r[k].split("(")[0]
        tot[name] += val * scale
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {total / 1e3:.3f} ms of kernel time (ncu-serialised, cold cache: compare shares)")
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for name, t in tot.most_common():
        print(f"{name[:70]:70s} {cnt[name]:8d} {t:12.1f} {t / cnt[name]:10.2f} {100 * t / total:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
