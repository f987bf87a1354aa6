"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list -> text summary for profiles/.

    python tools/launch_summary.py gpurun_out/r01c_launches.csv profiles/r01c_launches_summary.txt "command line..."
"""
import csv
import sys
from collections import OrderedDict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu].replace("second", "s").replace("n", "n"), None)
        unit = r[iu]
        scale = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1e-6)
        name = r[ik].split("(")[0]
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + v * scale)
    total = sum(t for _, t in tot.values())
    lines = [note, "(per-launch times under ncu are cold-cache and serialised: compare SHARES)", ""]
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{name:60s} launches {n:4d}  total {t:10.2f} ms  share {100 * t / total:5.1f} %")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
