"""Condense `ncu --page raw --csv` / `--page details --csv` exports into the small tables committed under profiles/.
usage: python tools/ncu_extract.py raw <raw.csv> <out.csv>            key metrics of every captured launch
       python tools/ncu_extract.py launches <launches.csv> [name]     per-kernel totals and shares of a gpu__time_duration launch list
       python tools/ncu_extract.py details <details.csv> <out.csv> <launch ids...>   details-page rows of the chosen launches"""
import csv
import sys
from collections import defaultdict

KEYS = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def rows_of(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.reader(lines))


def raw(src, dst):
    r = rows_of(src)
    head, units, data = r[0], r[1], r[2:]
    idx = [head.index(k) for k in KEYS if k in head]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx]); w.writerow([units[i] for i in idx])
        for d in data:
            name = d[head.index("Kernel Name")].split("(")[0]
            w.writerow([name if head[i] == "Kernel Name" else d[i] for i in idx])


def launches(src, only=None):
    r = rows_of(src)
    head = r[0]
    ni, vi = head.index("Kernel Name"), head.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for d in r[1:]:
        if len(d) <= vi:
            continue
        name = d[ni].split("(")[0].split("<")[0]
        try:
            v = float(d[vi].replace(",", ""))
        except ValueError:
            continue
        unit = d[head.index("Metric Unit")]
        v = v / 1e3 if unit in ("ns", "nsecond") else v
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if only and only not in name:
            continue
        print(f"{name:40s} launches {cnt[name]:5d}  total {v:10.1f} us  share {100 * v / s:5.1f} %  avg {v / cnt[name]:8.2f} us")
    print(f"{'ALL':40s} launches {sum(cnt.values()):5d}  total {s:10.1f} us")


def details(src, dst, ids):
    r = rows_of(src)
    head = r[0]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(head)
        for d in r[1:]:
            if d[0] in ids:
                w.writerow(d)


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "raw":
        raw(sys.argv[2], sys.argv[3])
    elif cmd == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        details(sys.argv[2], sys.argv[3], set(sys.argv[4:]))
