#!/usr/bin/env python
"""profiles/rNN_k3_traffic.json from an ncu CSV of ONE config-2 step's aligner launches.

On the GPU box (after the same command has run plain):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:align_locate -s <launches of the warm-up steps> -c <launches of one step> --csv \
        --log-file gpurun_out/k3_traffic.csv python tools/profile_step.py 100000 2
Here:
    python tools/k3_traffic.py gpurun_out/k3_traffic.csv 100000 profiles/r02_k3_traffic.json

The file carries a digest of the kernel sources (bench.source_sha): bench.py quotes the traffic only while it matches.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    src, reads, dst = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    import bench
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iid = hdr.index("ID")
    per = {}
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        if r[im].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        if r[im] == "gpu__time_duration.sum":
            v *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1}[u]
        per.setdefault((r[iid], r[ik]), {})[r[im]] = v
    out = {"what": f"ncu dram bytes of the {len(per)} aligner launches (strip pass align_locate_nb_kernel<S> + full-band pass "
                   f"align_locate_kernel<S>, one per band class) of ONE config-2 step ({reads} reads), serialised by ncu",
           "reads": reads, "source_sha": bench.source_sha(),
           "dram_bytes_read": sum(p.get("dram__bytes_read.sum", 0) for p in per.values()),
           "dram_bytes_write": sum(p.get("dram__bytes_write.sum", 0) for p in per.values()),
           "serialized_kernel_seconds": sum(p.get("gpu__time_duration.sum", 0) for p in per.values()),
           "per_kernel": {f"{k[1]} #{k[0]}": v for k, v in per.items()}}
    out["dram_bytes_total"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(f"{dst}: {len(per)} launches, read {out['dram_bytes_read'] / 1e9:.1f} GB, write {out['dram_bytes_write'] / 1e9:.1f} GB, "
          f"{out['serialized_kernel_seconds'] * 1e3:.1f} ms serialised, sources {out['source_sha']}")


if __name__ == "__main__":
    main()
