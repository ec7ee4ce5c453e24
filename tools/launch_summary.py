"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/ helper)"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        out.append((re.sub(r"\(.*", "", row["Kernel Name"])[:60], us, row["Grid Size"]))
    return out


def main(path, detail=None):
    rows = load(path)
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for n, us, g in rows:
        tot[n] = tot.get(n, 0) + us
        cnt[n] += 1
    T = sum(tot.values())
    print(f"total {T/1e3:.3f} ms over {len(rows)} launches")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:16]:
        print(f"{v/1e3:9.3f} ms {100*v/T:5.1f}% n={cnt[k]:4d} avg={v/cnt[k]:8.1f} us  {k}")
    if detail:
        print([(round(us, 1), g) for n, us, g in rows if detail in n][:80])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
