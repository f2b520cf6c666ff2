"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) per kernel:
python tools/ncu_launches_summary.py X.csv [out.md] [title]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    src = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    title = sys.argv[3] if len(sys.argv) > 3 else src
    lines = [ln for ln in open(src) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", r[ix["Kernel Name"]])
        name = re.sub(r"[(<].*", "", name)
        ns = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        ms = ns * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        t, c = agg.get(name, (0.0, 0))
        agg[name] = (t + ms, c + 1)
    total = sum(t for t, _ in agg.values())
    text = [f"# {title}", "",
            "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's `kernel_share`, not absolutes.", "",
            "| kernel | launches | total ms | avg ms | share |", "|---|---|---|---|---|"]
    for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        text.append(f"| {name} | {c} | {t:.3f} | {t / c:.4f} | {100 * t / total:.1f}% |")
    text = "\n".join(text) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
