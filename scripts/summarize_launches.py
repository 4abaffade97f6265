"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, title=""):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"^void ", "", name)
        m = re.match(r"([A-Za-z0-9_:]+(?:<[0-9, ]+>)?)", name)
        short = m.group(1) if m else name[:40]
        if "spin_kernel" in name or short.startswith("cuda::"):
            continue  # torch.cuda._sleep (bench.py parks the GPU before its per-launch timing pass)
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            ns *= 1e3
        rows.append((short, ns))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    tot = sum(v[1] for v in agg.values())
    print("# %s" % (title or path))
    print()
    print("%d launches, %.1f us total (cold-cache, serialised: compare SHARES)\n" % (len(rows), tot / 1e3))
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.0f | %.1f%% | %.1f |" % (k, n, ns / 1e3, 100 * ns / tot, ns / 1e3 / n))


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
