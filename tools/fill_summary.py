"""Fill the "Final state" table of profiles/r01_summary.md from the bench JSON files next to it."""
import json
import os
import re

here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def line(name):
    try:
        return json.load(open(os.path.join(here, name)))
    except (OSError, ValueError):
        return None


def cells(rec):
    if rec is None:
        return ["n/a"] * 4
    e2e = rec.get("e2e", {}).get("value")
    frac = (rec.get("roofline") or {}).get("frac")
    return [f"{rec['ms_per_step']:.3f}", f"{rec['value']:.1f}", f"{e2e:.1f}" if e2e else "n/a",
            f"{frac:.3f}" if frac else "n/a"]


def main():
    path = os.path.join(here, "r01_summary.md")
    s = open(path).read()
    out = []
    for ln in s.splitlines():
        m = re.match(r"\| `(r01_[a-z0-9_]+\.json)`(?: / `(r01_[a-z0-9_]+\.json)`)? \| ([^|]+) \|", ln)
        if m and "final" in m.group(1) or (m and "reference_arm" in m.group(1)):
            recs = [line(m.group(1))] + ([line(m.group(2))] if m.group(2) else [])
            vals = [cells(r) for r in recs]
            merged = [" / ".join(v[i] for v in vals) for i in range(4)]
            ln = f"| `{m.group(1)}`" + (f" / `{m.group(2)}`" if m.group(2) else "") + f" | {m.group(3).strip()} | " + " | ".join(merged) + " |"
        out.append(ln)
    open(path, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
