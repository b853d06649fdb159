"""bench.py -- seq/s of the OFASys unified encoder-decoder fwd+bwd hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): image_caption, OFA-base 12L/12L d=768, 224^2 patch-embed
(257 tokens) + 8-token prompt -> 64-token caption, bf16, per-GPU batch 64 ("B=32 (also 64)"; --batch 32 reproduces the smaller one), synthetic data, random-init
weights of that architecture.  A step = forward + sum-CE loss + backward of every parameter gradient
(+ the gradient all-reduce when N > 1); optimizer excluded.

One JSON line on stdout (rank 0):
  value      seq/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API with pinned HOST buffers: H2D of the step's inputs and
             D2H of the loss inside the timed region
  roofline   tcgen05 GEMM (dominant kernel): algorithmic FLOPs / per-launch CUDA-event time vs the
             measured bf16 peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle (CPU restatement of the reference, oracle/oracle_model.py) timed on host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "seq/s enc-dec fwd+bwd, OFA-base mixed-modality"
WORKLOAD = "image_caption OFA-base 12L/12L d=768, 224^2 patch-embed(257 tok)+8-tok prompt -> 64-tok caption"
V = 50265
CFG = dict(embed_dim=768, heads=12, ffn_dim=3072, enc_layers=12, dec_layers=12, vocab=V, mode="B")
PROMPT, TGT = 8, 64
FWD_GFLOP_PER_SEQ = 71.15  # SURVEY.md 8d (FlopCounterMode on the reference, cfg2a)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"], src="measured")
    return dict(tf_burst=1590.0, tf_sustained=1400.0, hbm=6650.0, src="fallback")


# ------------------------------------------------------------------------------------ model / data
def build_model(dev):
    import ofasys_b200 as ob

    cfg = ob.GeneralistModelConfig.default()
    cfg.dropout = 0.0
    cfg.use_self_attn_bias = False
    cfg.entangle_position_embedding = True
    cfg.arch = "base"
    m = ob.GeneralistModel(cfg)
    m.cfg.encoder.layers = m.cfg.decoder.layers = 12  # "OFA-base 12L/12L" (reference preset is 6/6; SURVEY.md header)
    for n in ("text", "image_patch_embed"):
        a = getattr(m.cfg.adaptor, n)
        a.is_active = True
        a.entangle_position_embedding = True
    m.cfg.adaptor.image_patch_embed.embed_dim = 768
    torch.manual_seed(0)
    m.initialize(ob.Dictionary(n_dummy=V - 4))
    return m.to(torch.bfloat16).to(dev).train()


def host_batch(B, seed, pin):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 3, 224, 224, generator=g)
    prompt = torch.randint(4, V, (B, PROMPT), generator=g)
    prev = torch.randint(4, V, (B, TGT), generator=g)
    prev[:, 0] = 0
    n_short = max(1, B // 10)  # 10 % of the sequences shortened and right-padded
    prev[:n_short, TGT - TGT // 4:] = 1
    tgt = torch.roll(prev, -1, 1)
    tgt[:, -1] = 2
    tgt[prev == 1] = 1
    tgt[torch.roll(prev == 1, -1, 1)] = 1
    out = dict(img=img, prompt=prompt, prev=prev, tgt=tgt)
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def to_slots(b):
    import ofasys_b200 as ob

    MT = ob.ModalityType
    return [ob.Slot(MT.IMAGE, True, b["img"], attributes="adaptor=image_patch_embed"), ob.Slot(MT.TEXT, True, b["prompt"]),
            ob.Slot(MT.TEXT, False, b["prev"])]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []
        self.stop = False
        self.index = index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in r.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(float(r[0])) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arms
def cpu_oracle_run(steps, warmup, batch, threads):
    """fwd+bwd of the oracle (reference algorithm, fp32) on host cores: bounded sample of the workload."""
    from oracle import cases
    from oracle import oracle_model as om

    torch.set_num_threads(threads)
    cfg = om.OracleConfig(**CFG)
    spec = _spec_cache()
    sd = cases.synth_state_dict(spec, seed=0)
    hb = host_batch(batch, 1234, pin=False)
    slots = [om.OSlot(om.IMAGE, True, hb["img"], adaptor="image_patch_embed"), om.OSlot(om.TEXT, True, hb["prompt"]), om.OSlot(om.TEXT, False, hb["prev"])]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        om.loss_and_grads(sd, cfg, slots, hb["tgt"])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return batch / med, med


def _spec_cache():
    """parameter name -> shape of the benchmark model (CPU construction only; no kernels run)."""
    import ofasys_b200 as ob
    import ofasys_b200.model.ofa as ofa_mod

    cfg = ob.GeneralistModelConfig.default()
    cfg.dropout = 0.0
    cfg.use_self_attn_bias = False
    cfg.entangle_position_embedding = True
    cfg.arch = "base"
    m = ob.GeneralistModel(cfg)
    m.cfg.encoder.layers = m.cfg.decoder.layers = 12
    for n in ("text", "image_patch_embed"):
        a = getattr(m.cfg.adaptor, n)
        a.is_active = True
        a.entangle_position_embedding = True
    m.cfg.adaptor.image_patch_embed.embed_dim = 768
    orig = ofa_mod.init_bert_params
    try:
        ofa_mod.init_bert_params = lambda mod: None  # shapes only
        m.initialize(ob.Dictionary(n_dummy=V - 4))
    finally:
        ofa_mod.init_bert_params = orig
    return {k: tuple(v.shape) for k, v in m.state_dict().items() if v.is_floating_point() and not k.endswith(".version")}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 32)  # more threads slow torch's CPU kernels down on these shapes
    batch = 4
    sps, med = cpu_oracle_run(max(1, min(args.steps, 3)), min(args.warmup, 1), batch, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "seq/s", "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)),
        "warmup": min(args.warmup, 1), "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": batch, "parallelism": "cpu"},
        "cpu_baseline": {"value": sps, "unit": "seq/s", "cores": threads, "kind": "port",
                         "sample": f"batch {batch} fwd+bwd, fp32, torch {torch.__version__} CPU, median of {max(1, min(args.steps, 3))} steps"},
        "e2e": {"value": sps, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rank, world, local_rank):
    import torch.distributed as dist

    from ofasys_b200 import _lib

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.check(_lib.lib().ofab_device_check(local_rank), "device check")
    if args.no_pdl:
        _lib.lib().ofab_set_pdl(0)
    if world > 1:
        # 64 MB gradient buckets: the Simple protocol (LL / LL128 trade bandwidth for latency); measured at N=2: 28.17 vs 28.67 ms
        os.environ.setdefault("NCCL_PROTO", "Simple")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    model = build_model(dev)
    params = [p for p in model.parameters() if p.requires_grad]
    hb = host_batch(B, 1234 + rank, pin=True)
    db = {k: v.to(dev) for k, v in hb.items()}  # static device inputs (graph replays read these addresses)
    ntok = int((hb["tgt"] != 1).sum())
    state = {"loss": None, "graph": None}

    def fwd_bwd():
        for p in params:
            p.grad = None
        loss = model.forward_loss(to_slots(db), db["tgt"])
        loss.backward()
        return loss

    from ofasys_b200.distributed import GradArena, GradBuckets

    # N > 1, default (--dp arena): the weight-gradient GEMMs write straight into one flat gradient arena (no pack / unpack
    # copies); it is cut into 64 MB buckets in backward order and every bucket is all-reduced (NCCL over NVLink, side stream)
    # as soon as its last gradient exists, while the rest of the backward runs -- the whole step incl. the collectives is ONE
    # CUDA graph.  --dp split: the older scheme (backward cut at the encoder/decoder boundary, flat buckets packed by a
    # multi-tensor copy).  N == 1: one graph, no exchange.
    if world > 1 and args.dp == "arena":
        fwd_bwd()  # the first forward packs the q|k|v parameter storages: the arena keys its slots by parameter address
        torch.cuda.synchronize()
    arena = GradArena(params) if (world > 1 and args.dp == "arena") else None
    split = world > 1 and args.dp == "split"
    side = torch.cuda.Stream()
    fwd_bwd_local = fwd_bwd  # no collectives: what the rank-0-only roofline / breakdown sections run
    if arena is not None:
        def fwd_bwd():  # noqa: F811  (the data-parallel step: exchange included)
            arena.begin_step(arm=True)
            loss = model.forward_loss(to_slots(db), db["tgt"])
            loss.backward()
            arena.finish()
            return loss

    def fwd_bwd_begin():
        for p in params:
            p.grad = None
        loss, early, finish = model.forward_backward_split(to_slots(db), db["tgt"])
        state["early"], state["finish"] = early, finish
        return loss

    use_graph = not args.no_graph
    if use_graph:
        # CUDA graph(s) of the whole fwd+bwd: ~1000 kernel launches per step are replayed without host work
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                if split:
                    fwd_bwd_begin()
                    state["finish"]()
                else:
                    fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

        def capture():
            c0 = _lib.launch_count
            g = torch.cuda.CUDAGraph()
            for p in params:
                p.grad = None
            if split:
                with torch.cuda.graph(g):
                    state["loss"] = fwd_bwd_begin()
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=g.pool()):
                    state["finish"]()
                state["graph2"] = g2
            else:
                with torch.cuda.graph(g):
                    state["loss"] = fwd_bwd()
            state["graph"] = g
            state["launches"] = _lib.launch_count - c0

        try:
            capture()
        except Exception as ex:
            print(f"[bench] CUDA graph capture failed ({type(ex).__name__}: {ex})", file=sys.stderr)
            torch.cuda.synchronize()
            use_graph = False
            if not args.no_pdl:  # programmatic edges are the newest piece of the capture: retry once without them
                _lib.lib().ofab_set_pdl(0)
                args.no_pdl = True
                try:
                    capture()
                    use_graph = True
                    print("[bench] captured without PDL", file=sys.stderr)
                except Exception as ex2:
                    print(f"[bench] capture without PDL failed too ({type(ex2).__name__}: {ex2}); running eagerly", file=sys.stderr)
                    torch.cuda.synchronize()

    buckets = early_b = late_b = None

    def step_resident():
        nonlocal buckets, early_b, late_b
        if not split:
            if use_graph:
                state["graph"].replay()
                loss = state["loss"]
            else:
                loss = fwd_bwd()
            if world > 1 and arena is None:  # unoverlapped exchange (--dp plain)
                if buckets is None:
                    buckets = GradBuckets(params)
                buckets.allreduce()
            return loss
        if use_graph:
            state["graph"].replay()
            loss = state["loss"]
        else:
            loss = fwd_bwd_begin()
        if early_b is None:
            ids = {id(p) for p in state["early"]}
            early_b = GradBuckets(state["early"])
            late_b = GradBuckets([p for p in params if id(p) not in ids])
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            early_b.allreduce()
        if use_graph:
            state["graph2"].replay()
        else:
            state["finish"]()
        late_b.allreduce()
        cur.wait_stream(side)
        return loss

    # End to end: every step's inputs cross PCIe from pinned host memory (h2d_bytes_per_step) and the loss is read back.
    # The copy of step k+1's batch runs on a copy stream into a staging set while step k computes (what a data loader
    # with a prefetch queue does); at the start of a step the staged batch moves into the graph's static input buffers
    # with a device-to-device copy once the copy stream's event has fired.
    copy_stream = torch.cuda.Stream()
    stage = {k: torch.empty_like(v) for k, v in db.items()}
    staged = {"ev": None}

    def prefetch_next():
        with torch.cuda.stream(copy_stream):
            for k in stage:
                stage[k].copy_(hb[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged["ev"] = ev

    def step_e2e():
        if staged["ev"] is None:
            prefetch_next()  # first step: nothing staged yet, its copy is exposed
        cur = torch.cuda.current_stream()
        cur.wait_event(staged["ev"])
        for k in db:
            db[k].copy_(stage[k], non_blocking=True)
        ev_used = torch.cuda.Event()
        ev_used.record(cur)
        copy_stream.wait_event(ev_used)  # the staging set is free again once the D2D copies have run
        loss = step_resident()
        prefetch_next()  # next step's H2D overlaps this step's kernels
        return loss.item()  # D2H read of the step result

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = _lib.launch_count
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_l = (_lib.launch_count - c0) // steps + (state.get("launches") if use_graph else 0)  # graph nodes + eager pack/unpack
        return t.item() / steps, n_l

    with ClockSampler(local_rank) as cs:
        ms_step, launches = timed(step_resident, args.steps, args.warmup)
    clocks = cs.summary()
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))

    # ---- roofline of the dominant kernel (tcgen05 GEMM)
    # The step's GEMM launch list (every ofab_gemm_bf16 / _splitk call of one fwd+bwd, same shapes, layouts, epilogues,
    # in order) is replayed back to back as ONE CUDA graph on the launch stream and bracketed by CUDA events:
    # achieved = sum of 2*M*N*K over the list / replay time.  (Operands are scratch tensors of the same shapes; three
    # rotating sets so consecutive launches do not hit the same lines.)  For reference the eager per-launch figure
    # (an event pair around every launch of a real step, which also times the host-side launch gaps) is kept as
    # `achieved_eager_events`.
    roof = None
    if rank == 0:
        pk = peaks()
        recs = []
        orig = _lib.call
        GEMMS = ("ofab_gemm_bf16", "ofab_gemm_bf16_splitk")

        def call(name, *a):
            if name not in GEMMS:
                return orig(name, *a)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(name, *a)
            e.record()
            if name == "ofab_gemm_bf16":  # M N K A lda a_mn B ldb b_mn bias residual ldr D ldd d_dt stream
                spec = (a[0], a[1], a[2], a[4], a[5], a[7], a[8], a[9] is not None and a[9].value is not None,
                        a[10] is not None and a[10].value is not None, a[11], a[13], a[14], False)
            else:  # M N K A lda a_mn B ldb b_mn D ldd d_dt ws ws_elems stream
                spec = (a[0], a[1], a[2], a[4], a[5], a[7], a[8], False, False, 0, a[10], a[11], True)
            recs.append((2.0 * a[0] * a[1] * a[2], s, e, spec))

        if arena is not None:
            arena.enabled = False  # the replays below are local (no collectives, no arena slots)
        _lib.call = call
        try:
            for _ in range(2):
                fwd_bwd_local()
            torch.cuda.synchronize()
            recs.clear()
            n_eager = max(2, min(args.steps, 5))
            for _ in range(n_eager):
                fwd_bwd_local()
            torch.cuda.synchronize()
        finally:
            _lib.call = orig
        fl = sum(r[0] for r in recs)
        tm = sum(r[1].elapsed_time(r[2]) for r in recs) * 1e-3
        ach_eager = fl / tm / 1e12
        launch_list = [r[3] for r in recs[: len(recs) // n_eager]]
        from ofasys_b200 import ops as _ops

        pool = {}

        def scratch(key, shape, dtype, idx):
            k = (key, shape, dtype, idx % 3)
            if k not in pool:
                pool[k] = (torch.randn(shape, device=dev) * 0.05).to(dtype) if dtype != torch.float32 else torch.zeros(shape, device=dev)
            return pool[k]

        def replay_list():
            for i, (M, N, K, lda, a_mn, ldb, b_mn, has_bias, has_res, ldr, ldd, d_dt, split) in enumerate(launch_list):
                A = scratch("A", (K if a_mn else M, lda), torch.bfloat16, i)    # same leading dimensions as the real call
                Bm = scratch("B", (K if b_mn else N, ldb), torch.bfloat16, i)
                odt = torch.float32 if d_dt == 0 else torch.bfloat16
                D = scratch("D", (M, ldd), odt, i)
                if split:
                    _ops.gemm_splitk(M, N, K, A, lda, a_mn, Bm, ldb, b_mn, D, ldd)
                else:
                    bias = scratch("bias", ((N + 7) // 8 * 8,), torch.bfloat16, i) if has_bias else None
                    res = scratch("res", (M, ldr), torch.float32, i) if has_res else None
                    _ops.gemm(M, N, K, A, lda, a_mn, Bm, ldb, b_mn, D, ldd, bias=bias, residual=res, ldr=ldr)

        fl_list = sum(2.0 * sp[0] * sp[1] * sp[2] for sp in launch_list)  # (the padded tied-logits width is counted as launched)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            replay_list()
            replay_list()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            replay_list()
        for _ in range(3):
            gg.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = 10
        e0.record()
        for _ in range(nrep):
            gg.replay()
        e1.record()
        torch.cuda.synchronize()
        t_list = e0.elapsed_time(e1) * 1e-3 / nrep
        ach = fl_list / t_list / 1e12
        del gg
        pool.clear()
        # DRAM traffic of the dominant kernel: measured once under ncu (dram__bytes_read.sum + dram__bytes_write.sum over the
        # GEMM launches of one step, tools/launch_list_summary.py -> profiles/r01_gemm_traffic.json), per launch like `achieved`
        # (profiles/r02_gemm_traffic.json, written by tools/launch_list_summary.py; only valid for the build it was captured on:
        # the file carries the hash of the kernel sources -- a stale capture is reported as null, not as evidence)
        traffic = traffic_note = None
        tp = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                from ofasys_b200.build import _stamp

                if int(tj.get("per_gpu_batch", -1)) == B and tj.get("build_stamp") == _stamp():
                    traffic, traffic_note = tj["dram_bytes_per_launch"], tj.get("note")
                else:
                    traffic_note = "profiles/r02_gemm_traffic.json was captured on a different build or batch: not reported"
            except Exception:
                pass
        roof = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05)", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["tf_sustained"], "traffic": traffic, "traffic_unit": "bytes/launch (ncu)", "traffic_note": traffic_note, "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']})",
                "how": "the step's GEMM launch list replayed back to back as one CUDA graph, CUDA events around the replay",
                "gemm_launches_per_step": len(launch_list), "gemm_ms_per_step": t_list * 1e3,
                "avg_launch_us": t_list * 1e6 / max(1, len(launch_list)),
                "achieved_eager_events": ach_eager, "gemm_ms_per_step_eager_events": tm * 1e3 / n_eager}

    # ---- the step that follows the measured path every update: gradient norm + clip + fp32-master
    # Adam over all parameters (csrc/optim.cu), timed on the gradients the last replay left in p.grad.  HBM-bound:
    # 2 B (norm pass) + 2 + 12 B read, 12 + 2 B written per element.  Reported next to the headline, not inside it
    # (the metric is fwd+bwd, optimizer step excluded: SURVEY 8d).
    optim_rec = None
    if rank == 0:
        try:
            from ofasys_b200 import FusedAdam

            gparams = [p for p in params if p.grad is not None]
            n_el = sum(p.numel() for p in gparams)
            opt = FusedAdam(gparams, lr=1e-12, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)

            def opt_step():
                opt._table_ready = False
                opt.multiply_grads(1.0 / max(1, ntok))
                opt.clip_grad_norm(1.0)
                opt.step()

            for _ in range(3):
                opt_step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                opt_step()
            e1.record()
            torch.cuda.synchronize()
            t_opt = e0.elapsed_time(e1) / 10
            pk = peaks()
            gbs = n_el * 30.0 / (t_opt * 1e-3) / 1e9
            optim_rec = {"kernels": "adam_sqnorm_kernel + adam_norm_final_kernel + adam_step_kernel", "ms_per_update": t_opt, "elements": n_el,
                         "algorithmic_bytes_per_element": 30, "achieved_gbs": gbs, "peak_gbs": pk["hbm"], "frac": gbs / pk["hbm"], "bound": "hbm"}
            del opt
        except Exception as ex:
            optim_rec = {"error": f"{type(ex).__name__}: {ex}"}

    if args.kprofile:  # every rank runs the steps (collectives), rank 0 writes
        # kernel-level timeline of the replayed step (CUPTI via torch.profiler): hot caches, real back-to-back execution
        from torch.profiler import ProfilerActivity, profile

        for _ in range(2):
            step_resident()
        torch.cuda.synchronize()
        nrep = 5
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(nrep):
                step_resident()
            torch.cuda.synchronize()
        agg = {}
        for evt in prof.events():
            if evt.device_type is not None and "cuda" in str(evt.device_type).lower():
                a = agg.setdefault(evt.name, [0, 0.0])
                a[0] += 1
                a[1] += evt.device_time if hasattr(evt, "device_time") else evt.cuda_time
        rows = sorted(((k, v[0] / nrep, v[1] / nrep / 1e3) for k, v in agg.items()), key=lambda r: -r[2])
        out = {"step_kernel_ms": sum(r[2] for r in rows), "kernels": [{"name": k[:160], "launches_per_step": n, "ms_per_step": ms} for k, n, ms in rows]}
        if world > 1:  # where the exchange kernels sit on the timeline (us from the first kernel of the profile)
            devs = [e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()]
            t0 = min(e.time_range.start for e in devs)
            out["span_ms_per_step"] = (max(e.time_range.end for e in devs) - t0) / nrep / 1e3
            out["exchange_timeline"] = [{"name": e.name[:60], "start_us": e.time_range.start - t0, "dur_us": e.time_range.end - e.time_range.start}
                                        for e in devs if ("nccl" in e.name.lower() or "multi_copy" in e.name)][:200]
        if rank == 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kprofile.json"), "w"), indent=1)

    if rank == 0 and args.breakdown:
        # per-entry-point device time (CUDA events around every C-ABI call): where the step goes
        recs = []
        orig = _lib.call

        def call_all(name, *a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(name, *a)
            e.record()
            recs.append((name, s, e))

        _lib.call = call_all
        try:
            fwd_bwd_local()
            torch.cuda.synchronize()
            recs.clear()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(3):
                fwd_bwd_local()
            t1.record()
            torch.cuda.synchronize()
        finally:
            _lib.call = orig
        agg = {}
        for name, s, e in recs:
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += s.elapsed_time(e)
        tot = t0.elapsed_time(t1) / 3
        out = {k: {"calls_per_step": v[0] / 3, "ms_per_step": v[1] / 3} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        out["_sum_ms"] = sum(v[1] for v in agg.values()) / 3
        out["_step_ms_with_events"] = tot
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "breakdown.json"), "w"), indent=1)

    # ---- the other BASELINE.json workloads (configs[2..4]: ASR, co-training, OFA-large video + grounding), same method, at
    # this N -- every rank runs them (their steps end with the gradient exchange); see workloads.py
    extra, dp_check = [], None
    peak_mem_gb = torch.cuda.max_memory_allocated(dev) / 1e9  # of the headline workload (before the other workloads allocate)
    if not args.no_workloads:
        import workloads as wl

        state.clear()
        if arena is not None:
            arena.close()
        del model, params
        torch.cuda.empty_cache()
        pk = peaks()
        for nm in args.workloads.split(","):
            try:
                r = wl.run_workload(nm, dev, steps=max(3, min(args.steps, 5)), warmup=3, world=world)
                r["model_frac_of_bf16_peak"] = r["model_tflops_per_gpu"] / pk["tf_sustained"]
            except Exception as ex:
                r = {"workload": nm, "error": f"{type(ex).__name__}: {str(ex)[:300]}"}
                torch.cuda.synchronize()
            extra.append(r)
        try:  # model-level data-parallel gate: N-rank gradients == one process on the concatenated batch
            dp_check = wl.dp_equality_check(dev, world, rank)
        except Exception as ex:
            dp_check = {"ok": False, "error": f"{type(ex).__name__}: {str(ex)[:300]}"}

    if rank == 0:
        cpu = None
        if not args.no_cpu:
            threads = min(os.cpu_count() or 1, 32)  # more threads slow torch's CPU kernels down on these shapes
            try:
                sps, med = cpu_oracle_run(2, 1, 4, threads)
                cpu = {"value": sps, "unit": "seq/s", "cores": threads, "kind": "port",
                       "sample": "oracle (CPU restatement of the reference), batch 4 fwd+bwd fp32, median of 2 steps after 1 warm-up"}
            except Exception as ex:  # never hide the GPU number behind a CPU-side failure
                cpu = {"value": None, "unit": "seq/s", "cores": threads, "kind": "port", "sample": f"failed: {ex}"}
        gb = B * world
        value = gb / (ms_step * 1e-3)
        pk = peaks()
        h2d = sum(v.numel() * v.element_size() for v in hb.values())
        line = {
            "metric": METRIC, "value": value, "unit": "seq/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": gb, "per_gpu_batch": B, "src_len": 257 + PROMPT, "tgt_len": TGT,
                       "parallelism": f"dp{world}", "cuda_graph": bool(use_graph), "pdl": not args.no_pdl,
                       "grad_exchange": (None if world == 1 else {"arena": "gradient arena (GEMMs write dW in place), NCCL all-reduce(avg) per 64 MB bucket as soon as its last gradient exists, overlapped with backward, captured in the step graph", "split": "NCCL all-reduce(avg) of packed 64 MB buckets; decoder-side buckets overlap the encoder backward", "plain": "NCCL all-reduce(avg) of packed 64 MB buckets after backward"}[args.dp]), "l2": "working set per step (0.4 GB weights + >5 GB activations) >> 126 MB L2; no flush needed",
                       "algorithmic_gflop_per_seq": 3 * FWD_GFLOP_PER_SEQ, "ntokens_per_rank": ntok},
            "clocks": clocks,
            "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "seq/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e,
                    "how": "pinned-host batch -> H2D on a copy stream (prefetch of the next step's batch) -> D2D into the graph's input buffers -> fwd+bwd -> loss.item()"},
            "gpu_launches": launches,
            "peak_mem_gb": peak_mem_gb,
            "ce_chunk_rows": int(os.environ.get("OFAB_CE_CHUNK_ROWS", "0")) or None,
            "model_tflops": value * 3 * FWD_GFLOP_PER_SEQ / 1e3 / world,
            "model_frac_of_bf16_peak": value * 3 * FWD_GFLOP_PER_SEQ / 1e3 / world / pk["tf_sustained"],
            "roofline": roof,
            "optimizer_step": optim_rec,
            "cpu_baseline": cpu,
            "workloads": extra,
            "dp_check": dp_check,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing the communicator down: destroy_process_group() after NCCL kernels were captured into CUDA
        # graphs was observed to hang (both ranks idle after the line was printed).  Everything measured is on stdout.
        torch.cuda.synchronize()
        try:
            dist.barrier()
        except Exception:
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (SURVEY 8d: 32 or 64; measured 1.20x seq/s at 64 vs 32, 1.03x more at 128)")
    ap.add_argument("--impl", default="ofab", choices=["ofab", "reference"])
    ap.add_argument("--breakdown", action="store_true", help="also write gpurun_out/breakdown.json (per entry point device time)")
    ap.add_argument("--kprofile", action="store_true", help="also write gpurun_out/kprofile.json (per-kernel device time of the replayed step, CUPTI)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--dp", default="arena", choices=["arena", "split", "plain"], help="N>1 gradient exchange: arena (gradients written into a flat arena, per-bucket all-reduce overlapped with backward), split (backward cut at the encoder/decoder boundary, packed buckets), plain (all after backward)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-workloads", action="store_true", help="headline workload only (skip the configs[2..4] lines and the data-parallel gate)")
    ap.add_argument("--workloads", default="asr,cotrain,large", help="comma-separated extra workloads (workloads.py)")
    ap.add_argument("--no-pdl", action="store_true", help="launch without programmatic dependent launch (A/B of the kernel-boundary overlap)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
