"""Summarise a `bench.py --timeline` CSV: span, union-busy time, per-kernel in-situ durations, idle gaps."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [(float(r["start_us"]), float(r["dur_us"]), r["name"]) for r in csv.DictReader(open(path))]
    rows.sort()
    span = max(s + d for s, d, _ in rows) - rows[0][0]
    # union of busy intervals
    busy, cur_s, cur_e = 0.0, None, None
    gaps = []
    for s, d, _ in rows:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
                gaps.append(s - cur_e)
            cur_s, cur_e = s, s + d
        else:
            cur_e = max(cur_e, s + d)
    busy += cur_e - cur_s
    total = sum(d for _, d, _ in rows)
    print(f"# {path}: {len(rows)} kernels, span {span / 1e3:.2f} ms, union busy {busy / 1e3:.2f} ms ({busy / span * 100:.1f}%), "
          f"sum of durations {total / 1e3:.2f} ms (overlap factor {total / busy:.2f})")
    if gaps:
        gaps.sort()
        print(f"# idle gaps: n={len(gaps)} total={sum(gaps) / 1e3:.2f} ms median={gaps[len(gaps) // 2]:.2f} us p90={gaps[int(len(gaps) * .9)]:.2f} us")
    agg = defaultdict(lambda: [0, 0.0])
    for _, d, n in rows:
        k = n.split("(")[0].replace("void ", "").replace("reppo::", "")[:70]
        agg[k][0] += 1
        agg[k][1] += d
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{k:70s} n={n:6d} total={t / 1e3:8.2f} ms avg={t / n:7.2f} us share_of_sum={t / total * 100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
