import json, re, sys, collections
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/kprofile.json"))
agg = collections.OrderedDict()
for k in d["kernels"]:
    n = k["name"]
    m = re.search(r"gemm_bf16_kernel<(.*?)>", n)
    if m:
        key = "gemm<" + m.group(1) + ">"
    else:
        key = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        key = re.sub(r"\(.*", "", key)
    a = agg.setdefault(key[:100], [0, 0.0]); a[0] += k["launches_per_step"]; a[1] += k["ms_per_step"]
print(f"sum of kernel time per step: {d['step_kernel_ms']:.2f} ms")
fam = collections.defaultdict(float)
for k, v in agg.items():
    f = "gemm" if k.startswith("gemm") else "attention" if "attn" in k else "layernorm" if k.startswith("ln_") or "reduce_partials" in k else "colsum" if "colsum" in k else "torch" if k.startswith("at::") else "other"
    fam[f] += v[1]
print("  ".join(f"{f}={t:.2f}ms" for f, t in sorted(fam.items(), key=lambda kv: -kv[1])))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1]:7.3f} ms {v[0]:7.1f} x {v[1]/max(v[0],1e-9)*1e3:7.1f} us  {k}")
