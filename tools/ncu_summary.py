"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed): python tools/ncu_summary.py rep.ncu-rep [out.md]"""
import collections
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_%"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"), ("l1tex__t_sector_hit_rate.pct", "l1hit_%"),
        ("lts__t_sector_hit_rate.pct", "l2hit_%"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(unit, 1)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        d = agg.setdefault(name, collections.defaultdict(list))
        for m, short in WANT:
            if m not in col:
                continue
            v, u = r[col[m]], units[col[m]]
            if short == "time":
                d[short].append(to_us(v, u))
            elif short in ("dram_rd", "dram_wr"):
                d[short].append(to_bytes(v, u))
            else:
                d[short].append(float(v.replace(",", "")))
    lines = ["| kernel | launches | time µs | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | sm % | fp64 % | regs | occ % | L1 hit % | L2 hit % | grid×block |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    out = {}
    for k, d in agg.items():
        n = len(d["time"])
        av = lambda key: sum(d[key]) / len(d[key]) if d[key] else float("nan")
        t, rd, wr = av("time"), av("dram_rd"), av("dram_wr")
        lines.append("| %s | %d | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %d | %.1f | %.1f | %.1f | %d×%d |" % (
            k, n, t, rd / 1e6, wr / 1e6, (rd + wr) / t / 1e3, av("dram_%"), av("sm_%"), av("fp64_%"), av("regs"), av("occ_%"), av("l1hit_%"), av("l2hit_%"),
            av("grid"), av("block")))
        out[k] = {"time_us": t, "dram_bytes_per_launch": rd + wr}
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(text + "\n")
    return out


if __name__ == "__main__":
    main()
