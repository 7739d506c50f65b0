"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python tools/launch_list_summary.py in.csv "title" > out.md"""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        ns = float(r[col["Metric Value"]].replace(",", ""))
        if r[col["Metric Unit"]] in ("us", "usecond"):
            ns *= 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print("# " + (sys.argv[2] if len(sys.argv) > 2 else "launch list"))
    print("\nCold-cache, serialised per-launch times (ncu replays every kernel in isolation): compare SHARES, not absolutes.\n")
    print("| kernel | launches | total µs | avg µs | share of all launches |")
    print("|---|---|---|---|---|")
    for k, (n, ns) in agg.items():
        print("| %s | %d | %.1f | %.1f | %.1f %% |" % (k, n, ns / 1e3, ns / 1e3 / n, 100 * ns / tot))
    print("\ntotal %.1f µs over %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))


if __name__ == "__main__":
    main()
