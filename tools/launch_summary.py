"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel family:
count, total time, share of the listed launches.   usage: python tools/launch_summary.py launches.csv > summary.txt"""
import csv
import re
import sys
from collections import defaultdict


def family(name: str) -> str:
    m = re.search(r"(gemm_tc_kernel<[^>]*>+|attn_fwd_kernel<[^>]*>|\w+_kernel\b|\w+Kernel\w*|cutlass\w*|cudnn\w*|sm\d+_\w+|nvjet\w*)", name)
    short = m.group(1) if m else name
    short = re.sub(r"mvoc::(gemm::)?", "", short)
    return short[:90]


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    i_name, i_val, i_metric, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    n = 0
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        us = v / 1e3 if r[i_unit] in ("ns", "nsecond") else (v if r[i_unit] in ("us", "usecond") else v * 1e3)
        a = agg[family(r[i_name])]
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    print(f"# {path}: {n} launches, {total / 1e3:.2f} ms summed (cold-cache, serialised under ncu: compare SHARES)")
    print(f"{'kernel family':92s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:92s} {c:8d} {us / 1e3:9.3f} {100 * us / total:6.1f}%")


if __name__ == "__main__":
    main()
