"""The BASELINE.json workloads as product-API models + synthetic batches (bench.py and tests/ share this; no oracle/).

BASELINE.json `configs` (sizes as SURVEY.md 8d prescribes them):
  caption   configs[1]  image_caption, OFA-base 12L/12L d=768, 224^2 patch-embed (257 tok) + 8-tok prompt -> 64-tok caption  (Mode B)
  asr       configs[2]  ASR, OFA-base, fbank [B, 998, 80] (10 s @ 16 kHz, ragged) + 12-tok prompt -> 128-tok transcript      (Mode A)
  cotrain   configs[3]  caption (ResNet-101 224^2, B=16) + VQA (16-tok question -> 8-tok answer, B=16) + text_infilling
                        (128 -> 128, B=16) per GPU, gradients accumulated over the three tasks, ONE exchange per step
                        (reference: engine/trainer.py:747-830)                                                                (Mode A)
  large     configs[4]  OFA-large 24L/12L d=1024 H=16: video_caption (16 x 224^2 frames -> S = 3136 + 8, T = 64, B=2) +
                        visual_grounding (512^2 image -> S = 1024 + 16, BOX target T = 5, B=8), ResNet-152, V = 50265 + 1000
                        `<bin>` tokens (reference: model/ofa.py:604-610, adaptor/video_image_sequence.py:111-208,
                        adaptor/image_resnet.py:116-202)                                                                      (Mode A)

A "step" of a workload = for every task batch: forward + sum-CE + backward, gradients accumulating in .grad; seq/s counts
the sequences of all its task batches.  FLOP figures are SURVEY 8d's 3 x F_fwd per sequence (FlopCounterMode on the
reference).
"""
import torch

V_TEXT = 50265
N_BINS = 1000

# per task: kind, per-GPU batch, source text length, target length, algorithmic GFLOP per sequence (3 x F_fwd)
WORKLOADS = {
    "caption": dict(arch="base", mode="B", adaptors=("text", "image_patch_embed"), vocab=V_TEXT,
                    tasks=[dict(name="image_caption", kind="patch", B=64, S=8, T=64, gflop=213.45)]),
    "asr": dict(arch="base", mode="A", adaptors=("text", "audio_fbank"), vocab=V_TEXT,
                tasks=[dict(name="asr", kind="audio", B=32, S=12, T=128, L=998, gflop=445.52)]),
    "cotrain": dict(arch="base", mode="A", adaptors=("text", "image_resnet"), vocab=V_TEXT, resnet_type="resnet101",
                    tasks=[dict(name="image_caption", kind="resnet", B=16, S=8, T=64, image=224, gflop=226.64),
                           dict(name="vqa", kind="resnet", B=16, S=16, T=8, image=224, gflop=183.23),
                           dict(name="text_infilling", kind="text", B=16, S=128, T=128, gflop=190.23)]),
    "large": dict(arch="large", mode="A", adaptors=("text", "image_resnet", "video_image_sequence"), vocab=V_TEXT + N_BINS,
                  resnet_type="resnet152",
                  tasks=[dict(name="video_caption", kind="video", B=2, S=8, T=64, image=224, frames=16, gflop=10374.0),
                         dict(name="visual_grounding", kind="resnet", B=8, S=16, T=5, image=512, box=True, gflop=2736.4)]),
}
ARCH = {"base": dict(d=768, heads=12, ffn=3072, enc=12, dec=12), "large": dict(d=1024, heads=16, ffn=4096, enc=24, dec=12)}


def describe(name):
    w = WORKLOADS[name]
    a = ARCH[w["arch"]]
    t = " + ".join(f"{t['name']} B={t['B']}" for t in w["tasks"])
    return f"{name}: OFA-{w['arch']} {a['enc']}L/{a['dec']}L d={a['d']} mode {w['mode']}; per-GPU step = {t}"


def build_model(name, dev, dtype=torch.bfloat16, seed=0, layers=None):
    """The ofasys_b200 GeneralistModel of workload `name` with random-init weights of that architecture."""
    import ofasys_b200 as ob

    w = WORKLOADS[name]
    a = ARCH[w["arch"]]
    cfg = ob.GeneralistModelConfig.default()
    cfg.dropout = 0.0
    cfg.attention_dropout = 0.0
    mode_b = w["mode"] == "B"
    if mode_b:
        cfg.use_self_attn_bias = False
        cfg.entangle_position_embedding = True
    m = ob.GeneralistModel(cfg)
    d = a["d"]
    m.cfg.encoder.embed_dim = m.cfg.decoder.embed_dim = d
    m.cfg.encoder.ffn_embed_dim = m.cfg.decoder.ffn_embed_dim = a["ffn"]
    m.cfg.decoder.input_dim = m.cfg.decoder.output_dim = d
    m.cfg.encoder.attention_heads = m.cfg.decoder.attention_heads = a["heads"]
    m.cfg.encoder.layers, m.cfg.decoder.layers = layers or (a["enc"], a["dec"])
    for ad in w["adaptors"]:
        acfg = getattr(m.cfg.adaptor, ad)
        acfg.is_active = True
        if mode_b:
            acfg.entangle_position_embedding = True
        if ad == "image_patch_embed":
            acfg.embed_dim = d
        if ad == "image_resnet":
            acfg.resnet_type = w["resnet_type"]
    if "resnet_type" in w:
        m.cfg.adaptor.image_resnet.resnet_type = w["resnet_type"]
    torch.manual_seed(seed)
    m.initialize(ob.Dictionary(n_dummy=V_TEXT - 4, num_bins=w["vocab"] - V_TEXT))
    return m.to(dtype).to(dev).train()


def _prev_target(g, B, T, vocab_hi, lo=4):
    prev = torch.randint(lo, vocab_hi, (B, T), generator=g)
    prev[:, 0] = 0
    n_short = max(1, B // 10)  # 10 % of the sequences shortened and right-padded
    if T >= 8:
        prev[:n_short, T - T // 4:] = 1
    tgt = torch.roll(prev, -1, 1)
    tgt[:, -1] = 2
    tgt[prev == 1] = 1
    tgt[torch.roll(prev == 1, -1, 1)] = 1
    return prev, tgt


def host_batch(task, vocab, seed, pin=False, B=None):
    """One synthetic task batch on the host: dict of tensors (inputs + `prev` + `tgt`)."""
    g = torch.Generator().manual_seed(seed)
    B = B or task["B"]
    out = {}
    k = task["kind"]
    if k == "patch" or k == "resnet":
        out["img"] = torch.randn(B, 3, task.get("image", 224), task.get("image", 224), generator=g)
    elif k == "video":
        out["video"] = torch.randn(B, 3, task["frames"], task["image"], task["image"], generator=g)
    elif k == "audio":
        out["fbank"] = torch.randn(B, task["L"], 80, generator=g)
        out["fbank_lengths"] = torch.randint(700, task["L"] + 1, (B,), generator=g)
    out["prompt"] = torch.randint(4, V_TEXT, (B, task["S"]), generator=g)
    if task.get("box"):  # bos + 4 `<bin>` tokens -> 4 bins + eos (preprocessor/default/box.py:101-110)
        bins = V_TEXT + torch.randint(0, N_BINS, (B, 4), generator=g)
        out["prev"] = torch.cat([torch.zeros(B, 1, dtype=torch.long), bins], 1)
        out["tgt"] = torch.cat([bins, torch.full((B, 1), 2, dtype=torch.long)], 1)
    else:
        out["prev"], out["tgt"] = _prev_target(g, B, task["T"], V_TEXT)
    if pin:
        out = {k2: v.pin_memory() for k2, v in out.items()}
    return out


def to_slots(task, b):
    import ofasys_b200 as ob

    MT = ob.ModalityType
    k = task["kind"]
    slots = []
    if k == "patch":
        slots.append(ob.Slot(MT.IMAGE, True, b["img"], attributes="adaptor=image_patch_embed"))
    elif k == "resnet":
        slots.append(ob.Slot(MT.IMAGE, True, b["img"], attributes="adaptor=image_resnet"))
    elif k == "video":
        slots.append(ob.Slot(MT.VIDEO, True, b["video"]))
    elif k == "audio":
        slots.append(ob.Slot(MT.AUDIO, True, {"fbank": b["fbank"], "fbank_lengths": b["fbank_lengths"]}))
    slots.append(ob.Slot(MT.TEXT, True, b["prompt"]))
    slots.append(ob.Slot(MT.BOX if task.get("box") else MT.TEXT, False, b["prev"]))
    return slots


def gflop_per_step(name, batch_scale=1.0):
    return sum(t["B"] * batch_scale * t["gflop"] for t in WORKLOADS[name]["tasks"])


def seqs_per_step(name):
    return sum(t["B"] for t in WORKLOADS[name]["tasks"])


def run_workload(name, dev, steps=5, warmup=3, world=1, use_graph=True, seed=1234, layers=None, kprofile=None):
    """Time `steps` steps of workload `name` on `dev` (inputs resident, CUDA events on the launch stream; with world > 1
    every step ends with the gradient average over ranks and the time is the max over ranks).  Returns a dict."""
    import torch.distributed as dist

    from ofasys_b200 import _lib
    from ofasys_b200.distributed import GradArena

    w = WORKLOADS[name]
    tasks = w["tasks"]
    model = build_model(name, dev, layers=layers)
    params = [p for p in model.parameters() if p.requires_grad]
    rank = dist.get_rank() if world > 1 else 0
    batches = [{k: v.to(dev) for k, v in host_batch(t, w["vocab"], seed + 17 * i + 1000 * rank).items()} for i, t in enumerate(tasks)]
    state = {}

    # world > 1: gradients land in the flat arena, buckets are all-reduced as the LAST task's backward fills them (the
    # exchange is part of the step and of its CUDA graph)
    arena = None

    def compute():
        if arena is not None:
            arena.begin_step()
        else:
            for p in params:
                p.grad = None
        losses = []
        for k, (t, b) in enumerate(zip(tasks, batches)):
            loss = model.forward_loss(to_slots(t, b), b["tgt"])
            if arena is not None and k == len(tasks) - 1:
                arena.arm()
            loss.backward()  # accumulates over the tasks of the step (trainer.py:752-830)
            losses.append(loss)
        if arena is not None:
            arena.finish()
        return losses

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        state["losses"] = compute()  # (packs the q|k|v parameter storages before the arena looks at their addresses)
        if world > 1:
            arena = GradArena(params)
        for _ in range(2):
            state["losses"] = compute()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    launches = None
    graph = None
    if use_graph:
        try:
            c0 = _lib.launch_count
            graph = torch.cuda.CUDAGraph()
            for p in params:
                p.grad = None
            with torch.cuda.graph(graph):
                state["losses"] = compute()
            launches = _lib.launch_count - c0
        except Exception as ex:  # report and fall back to eager launches (still our kernels)
            import sys

            print(f"[workloads] {name}: CUDA graph capture failed ({type(ex).__name__}: {ex}); eager", file=sys.stderr)
            torch.cuda.synchronize()
            graph = None
    def step():
        if graph is not None:
            graph.replay()
        else:
            state["losses"] = compute()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = _lib.launch_count
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if kprofile:  # per-kernel device time of the replayed step (CUPTI via torch.profiler) -> json file
        import json as _json

        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        agg = {}
        for evt in prof.events():
            if evt.device_type is not None and "cuda" in str(evt.device_type).lower():
                a = agg.setdefault(evt.name, [0, 0.0])
                a[0] += 1
                a[1] += evt.device_time if hasattr(evt, "device_time") else evt.cuda_time
        rows = sorted(((k, v[0] / 3, v[1] / 3 / 1e3) for k, v in agg.items()), key=lambda r: -r[2])
        _json.dump({"workload": name, "step_kernel_ms": sum(r[2] for r in rows),
                    "kernels": [{"name": k[:160], "launches_per_step": n, "ms_per_step": m_} for k, n, m_ in rows]}, open(kprofile, "w"), indent=1)
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    eager_l = (_lib.launch_count - c0) // steps
    seqs = seqs_per_step(name) * world
    tf = gflop_per_step(name) / (ms * 1e-3) / 1e3  # per GPU
    out = {"workload": name, "what": describe(name), "n_gpus": world, "ms_per_step": ms, "seq_per_s": seqs / (ms * 1e-3),
           "seqs_per_step": seqs, "cuda_graph": graph is not None, "gpu_launches": (launches or 0) + eager_l,
           "algorithmic_gflop_per_step_per_gpu": gflop_per_step(name), "model_tflops_per_gpu": tf,
           "losses": [float(x) for x in state["losses"]], "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    if arena is not None:
        arena.close()
    del graph, model, batches
    state.clear()
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":  # python workloads.py asr cotrain large   (1 GPU probe; bench.py runs them at every N)
    import json
    import sys
    import traceback

    dev = torch.device("cuda:0")
    kp = "--kprofile" in sys.argv
    for nm in [a for a in sys.argv[1:] if not a.startswith("--")] or ["asr", "cotrain", "large"]:
        try:
            r = run_workload(nm, dev, kprofile=f"gpurun_out/r02_kprofile_{nm}.json" if kp else None)
        except Exception as ex:
            traceback.print_exc()
            r = {"workload": nm, "error": f"{type(ex).__name__}: {ex}"}
            torch.cuda.synchronize()
        print(json.dumps(r), flush=True)


def dp_equality_check(dev, world, rank):
    """Data-parallel correctness gate (reference semantics engine/trainer.py:857-860): the gradients every rank
    holds after `all_reduce_grads()` and the trainer's `world_size / sum(ntokens)` rescale equal the gradients ONE process
    computes on the concatenated batch, divided by the total token count.  Text-only Mode A model with an audio adaptor that no
    batch touches (unused parameters contribute zeros).  Runs on every rank; returns {"rel_l2", "worst", "ok"}."""
    import torch.distributed as dist

    import ofasys_b200 as ob
    from ofasys_b200.distributed import DataParallelModel

    cfg = ob.GeneralistModelConfig.default()
    cfg.dropout = cfg.attention_dropout = 0.0
    m = ob.GeneralistModel(cfg)  # tiny: 4L/4L d=256 H=4
    for ad in ("text", "audio_fbank"):
        getattr(m.cfg.adaptor, ad).is_active = True
    torch.manual_seed(0)
    V = 1000
    m.initialize(ob.Dictionary(n_dummy=V - 4))
    m = m.to(torch.bfloat16).to(dev).train()
    model = DataParallelModel(m)
    MT = ob.ModalityType

    def shard(r):
        g = torch.Generator().manual_seed(4242 + r)
        B, S, T = 3, 20, 12
        src = torch.randint(4, V, (B, S), generator=g)
        src[-1, S - 5:] = 1
        prev, tgt = _prev_target(g, B, T, V)
        return src, prev, tgt

    def run(src, prev, tgt):
        for p in m.parameters():
            p.grad = None
        loss = model.module.forward_loss([ob.Slot(MT.TEXT, True, src.to(dev)), ob.Slot(MT.TEXT, False, prev.to(dev))], tgt.to(dev))
        loss.backward()
        return int((tgt != 1).sum())

    ntok = run(*shard(rank))
    model.all_reduce_grads()
    tot = torch.tensor([float(ntok)], device=dev)
    if world > 1:
        dist.all_reduce(tot)
    got = {k: (torch.zeros_like(p) if p.grad is None else p.grad.float() * (world / tot.item())) for k, p in m.named_parameters()}
    parts = [shard(r) for r in range(world)]
    n_all = run(*(torch.cat([p[i] for p in parts], 0) for i in range(3)))
    assert n_all == int(tot.item())
    num = den = 0.0
    worst = ("", 0.0)
    for k, p in m.named_parameters():
        want = torch.zeros_like(p).float() if p.grad is None else p.grad.float() / n_all
        num += (got[k] - want).pow(2).sum().item()
        den += want.pow(2).sum().item()
        nrm = want.norm().item()
        if nrm > 1e-3 * (den ** 0.5 + 1e-30):
            e = ((got[k] - want).norm() / nrm).item()
            if e > worst[1]:
                worst = (k, e)
    unused_zero = all(not got[k].any() for k in got if "audio_fbank" in k)
    e = (num / max(den, 1e-30)) ** 0.5
    for p in m.parameters():
        p.grad = None
    return {"what": f"{world}-rank gradients after all_reduce_grads() x world/sum(ntokens) vs one process on the concatenated batch",
            "rel_l2": e, "worst_param": worst[0], "worst_rel_l2": worst[1], "unused_adaptor_grads_zero": bool(unused_zero),
            "ok": bool(e <= 2e-2 and worst[1] <= 6e-2 and unused_zero)}
