"""Turn an `ncu --page raw --csv` export (with dram__bytes_read.sum / dram__bytes_write.sum) into
profiles/<name>_gemm_traffic.json: average DRAM bytes per edge-MLP GEMM launch (the kernels
bench.py's `roofline` object describes) plus a per-kernel table for profiles/r01_summary.md.
Also the whole capture's DRAM bytes ("step": the capture is one eager step of the head) for the HBM view of the
step in bench.py's `roofline.hbm_view`.
usage: python tools/ncu_traffic.py raw.csv out.json [shape share_support(0|1) precision]"""
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
    out["step"] = {
        "launches": sum(v[0] for v in per.values()),
        "dram_read_bytes": sum(v[1] for v in per.values()),
        "dram_write_bytes": sum(v[2] for v in per.values()),
        "kernel_time_us_serialised": sum(v[3] for v in per.values()),
        "what": "every library kernel of ONE eager step, each launch profiled alone with a cold L2: an upper bound of "
                "the replayed step's DRAM reads (concurrent wgrad / dgrad launches share dy and H through the L2); "
                "write-backs that leave the L2 after a kernel ended are not attributed",
    }
    if len(sys.argv) > 5:
        out["config"] = {"shape": sys.argv[3], "share_support": bool(int(sys.argv[4])), "precision": sys.argv[5]}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("bytes_per_launch", "launches")}))


if __name__ == "__main__":
    main()
