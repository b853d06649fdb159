"""Summarise an ncu launch list of one bench step (gpu__time_duration.sum + dram bytes per launch, CSV from
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) by kernel family and write
the GEMM's average DRAM traffic per launch to profiles/<tag>_gemm_traffic.json, stamped with the hash of the kernel sources the
capture ran on (bench.py reports it as roofline.traffic only when the stamp equals the current build's).

    python tools/launch_list_summary.py gpurun_out/launches.csv 64 r02 > profiles/r02_launches_summary.txt
"""
import csv
import json
import os
import sys

path, batch = sys.argv[1], int(sys.argv[2])
tag = sys.argv[3] if len(sys.argv) > 3 else "r02"
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
idi = h.index("ID")
per = {}
for r in rows[hdr + 1:]:
    d = per.setdefault(r[idi], {"name": r[ki]})
    v = float(r[vi].replace(",", "")) if r[vi] not in ("", "n/a") else 0.0
    u = r[ui]
    if "byte" in u.lower():
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if u == "us":
        v *= 1e3
    if u == "ms":
        v *= 1e6
    d[r[mi]] = v


def fam(n):
    if "gemm_bf16" in n:
        return "gemm (tcgen05)"
    if "attn_" in n:
        return "attention"
    if "embed_ln" in n:
        return "embed/CE"
    if "ln_" in n:
        return "LayerNorm family"
    if "ce_" in n:
        return "embed/CE"
    if "at::" in n or "elementwise" in n:
        return "torch elementwise / cat / fill"
    return "reductions / casts / copies"


agg = {}
for d in per.values():
    a = agg.setdefault(fam(d["name"]), [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"{len(per)} launches captured, {tot / 1e6:.3f} ms of kernel time (ncu: serialised, cold caches -- shares, not absolutes)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} launches {a[0]:5d}  time {a[1] / 1e6:8.3f} ms  share {a[1] / tot:6.1%}  dram {a[2] / 1e9:8.3f} GB  ({a[2] / max(a[0], 1) / 1e6:8.2f} MB/launch)")
g = agg.get("gemm (tcgen05)")
if g and g[0]:
    out = {"per_gpu_batch": batch, "dram_bytes_per_launch": g[2] / g[0], "gemm_launches": g[0], "gemm_time_share": g[1] / tot,
           "build_stamp": open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ofasys_b200", "build", "stamp")).read().strip(),
           "note": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the {g[0]} GEMM launches captured from one eager step at B={batch} ({os.path.basename(path)})"}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"{tag}_gemm_traffic.json"), "w"), indent=1)
