"""Run a few launches of one kernel family in isolation (for `ncu --set full` captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofasys_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
if which == "gemm":  # encoder fc1 forward of the benchmark step: M=8480 N=3072 K=768 (+bias)
    M, N, K = int(os.environ.get("GEMM_M", "8480")), 3072, 768
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for _ in range(5):
        ops.gemm(M, N, K, A, K, 0, B, K, 0, out, N, bias=bias)
elif which == "attn":  # encoder self-attention of the benchmark step: B=32 H=12 T=265
    B, T, H = 32, 265, 12
    qkv = torch.randn(B, T, 3 * H * 64, device=dev).bfloat16().requires_grad_(True)
    kpm = torch.zeros(B, T, dtype=torch.bool, device=dev)
    for _ in range(3):
        o = ops.attention(qkv, None, H, 0.125, None, kpm, False)
        o.backward(torch.randn_like(o))
elif which == "attn_modeA":  # encoder self-attention with OFA's position biases (ASR: S=260, 2047-bucket table), B=32 H=12
    B, T, H, NB = 32, 260, 12, 2047
    d = H * 64
    qkv = torch.randn(B, T, 3 * d, device=dev).bfloat16().requires_grad_(True)
    pq = torch.randn(1, T, d, device=dev).bfloat16().requires_grad_(True)
    pk = torch.randn(1, T, d, device=dev).bfloat16().requires_grad_(True)
    tab = torch.randn(NB, H, device=dev).bfloat16().requires_grad_(True)
    ar = torch.arange(T, device=dev)
    idx = ((ar[:, None] - ar[None, :]).clamp(-1023, 1023) + 1023).to(torch.int32)  # Toeplitz bucket map like the adaptors'
    idx[-12:, :] = -1
    idx[:, -12:] = -1  # the text prompt's block has its own ids; here: no relative bias across slots
    kpm = torch.zeros(B, T, dtype=torch.bool, device=dev)
    for _ in range(3):
        o = ops.attention(qkv, None, H, 0.0884, ops.PositionBias(pq, pk, idx.contiguous(), tab), kpm, False)
        o.backward(torch.randn_like(o))
elif which == "ln":  # GELU + ffn_layernorm of the benchmark step: rows 8480 x 3072
    rows = int(os.environ.get("LN_ROWS", "8480"))
    x = torch.randn(rows, 3072, device=dev).bfloat16().requires_grad_(True)
    w = torch.ones(3072, device=dev).bfloat16().requires_grad_(True)
    b = torch.zeros(3072, device=dev).bfloat16().requires_grad_(True)
    a = torch.randn(rows, 768, device=dev).bfloat16().requires_grad_(True)
    xr = torch.randn(rows, 768, device=dev).requires_grad_(True)
    ws = [(torch.rand(768, device=dev) + 0.5).bfloat16().requires_grad_(True) for _ in range(4)]
    for _ in range(3):
        y = ops.layer_norm(x, w, b, gelu=True)
        y.backward(torch.randn_like(y))
        xn, yy = ops.ln_res_ln(a, xr, ws[0], ws[1], ws[2], ws[3])  # the LN -> +residual -> LN junction
        torch.autograd.backward((xn, yy), (torch.randn_like(xn), torch.randn_like(yy)))
elif which == "adam":  # optimizer step over 200 M elements in 300 tensors (OFA-base sized)
    from ofasys_b200 import FusedAdam

    ps = [torch.nn.Parameter((torch.randn(768, 3072, device=dev) * 0.02).bfloat16()) for _ in range(80)] + \
         [torch.nn.Parameter((torch.randn(768, device=dev) * 0.02).bfloat16()) for _ in range(220)]
    for p in ps:
        p.grad = (torch.randn_like(p.float()) * 0.01).bfloat16()
    opt = FusedAdam(ps, lr=1e-4, weight_decay=0.01)
    for _ in range(3):
        opt._table_ready = False
        opt.multiply_grads(1.0 / 3000)
        opt.clip_grad_norm(1.0)
        opt.step()
torch.cuda.synchronize()
