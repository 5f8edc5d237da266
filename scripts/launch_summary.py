"""Per-kernel table from the ncu launch list of a bench.py run
(`ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file X.csv python bench.py ...`).
usage: python scripts/launch_summary.py gpurun_out/bench_launches.csv > profiles/rNN_bench_launches_summary.md"""
import csv
import re
import sys
from collections import OrderedDict


def main(path: str) -> None:
    lines = [l for l in open(path, newline="") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg: "OrderedDict[str, list[float]]" = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else v * (1000.0 if unit in ("ms", "msecond") else 1.0)
        agg.setdefault(r["Kernel Name"], []).append(us)
    ours = {k: v for k, v in agg.items() if any(ns in k for ns in ("trn::", "tc::", "attn::"))}
    total_ours = sum(sum(v) for v in ours.values())
    print("| kernel | launches | total us | mean us | share of our kernels |")
    print("|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        mine = k in ours
        label = re.sub(r"^void ", "", k).split("(")[0][:90]
        share = f"{100.0 * sum(v) / total_ours:.1f}%" if mine else "(torch helper: input generation)"
        print(f"| `{label if mine else 'torch:' + label[:60]}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.1f} | {share} |")
    pair = [v for k, v in ours.items() if "gemm_tf32x3_pair_kernel" in k]
    if pair:
        big = [x for x in pair[0] if x > 2000.0]
        small = [x for x in pair[0] if x <= 2000.0]
        pre = {k: v for k, v in ours.items() if "split_" in k}
        print()
        print(f"Resident 8192^3 steps of the CTA-pair kernel: {len(big)} launches, mean {sum(big) / max(len(big), 1):.1f} us; "
              f"row blocks of the host-slice path: {len(small)} launches, mean {sum(small) / max(len(small), 1):.1f} us.")
        for k, v in pre.items():
            bigv = [x for x in v if x > 60.0]
            if bigv:
                print(f"`{re.sub(r'^void ', '', k).split('(')[0]}` at full size: {len(bigv)} launches, mean {sum(bigv) / len(bigv):.1f} us.")


if __name__ == "__main__":
    main(sys.argv[1])
