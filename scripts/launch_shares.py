"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares."""
import collections
import csv
import sys


def main(path, out=None):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[start + 1:]:
        if len(r) <= mi:
            continue
        v = float(r[mi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].replace("octane::", "").replace("<unnamed>::", "")
        name = name.split("(")[0][:70]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    lines = [f"# {path}: {sum(cnt.values())} launches, {T / 1e3:.2f} ms of kernel time (ncu: cold-cache, serialised; shares only)"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        lines.append(f"{v / 1e3:10.3f} ms {100 * v / T:6.2f}%  n={cnt[k]:5d}  avg {v / cnt[k]:9.1f} us  {k}")
    s = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(s)
    print(s)


if __name__ == "__main__":
    main(*sys.argv[1:3])
