"""Turn an `ncu --page raw --csv` export (with dram__bytes_read.sum / dram__bytes_write.sum) into
profiles/<name>_gemm_traffic.json: average DRAM bytes per edge-MLP GEMM launch (the kernels
bench.py's `roofline` object describes) plus a per-kernel table for profiles/r01_summary.md.
usage: python tools/ncu_traffic.py raw.csv out.json"""
import collections
import csv
import json
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    per = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        for junk in ("void ", "mft::", "<unnamed>::", "(anonymous namespace)::"):
            name = name.replace(junk, "")
        name = name.split("(")[0]
        e = per[name]
        e[0] += 1
        e[1] += val(r, "dram__bytes_read.sum")
        e[2] += val(r, "dram__bytes_write.sum")
        e[3] += float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    gemm = {k: v for k, v in per.items() if k.startswith(("umma_rows_kernel", "umma_wgrad_kernel"))}
    n = sum(v[0] for v in gemm.values())
    total = sum(v[1] + v[2] for v in gemm.values())
    out = {
        "bytes_per_launch": total / n if n else None,
        "launches": n,
        "what": "mean of dram__bytes_read.sum + dram__bytes_write.sum over the umma_rows / umma_wgrad launches captured",
        "source": sys.argv[1],
        "per_kernel": {k: {"launches": v[0], "dram_read_MB": round(v[1] / v[0] / 1e6, 2),
                           "dram_write_MB": round(v[2] / v[0] / 1e6, 2), "avg_us": round(v[3] / v[0], 1)}
                       for k, v in sorted(per.items(), key=lambda kv: -kv[1][3])},
    }
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("bytes_per_launch", "launches")}))


if __name__ == "__main__":
    main()
