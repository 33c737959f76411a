"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals (markdown table).
usage: python tools/launch_summary.py gpurun_out/launches.csv [title]"""
import csv, re, sys
from collections import defaultdict

def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(float); cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("tsg::", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
        tot[name] += v_ms; cnt[name] += 1
    total = sum(tot.values())
    print(f"total kernel time {total:.1f} ms over {sum(cnt.values())} launches\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"| {k} | {cnt[k]} | {tot[k]:.2f} | {100 * tot[k] / total:.1f}% | {1e3 * tot[k] / cnt[k]:.1f} |")

if __name__ == "__main__":
    main()
