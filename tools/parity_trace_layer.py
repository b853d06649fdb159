"""Stage-by-stage comparison of ONE encoder layer of the CUDA path with the oracle's storage model (each oracle stage is fed the
CUDA path's own input of that stage): which stage leaves the model?"""
import os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cases, oracle_model as om
from ofasys_b200 import ops
from util import bf16_round_state_dict, build_product, load_golden, rel_l2, to_product_slots

name = sys.argv[1] if len(sys.argv) > 1 else "large_A"
li = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
g = load_golden(name)
sd = cases.synth_state_dict(g["spec"], seed=0)
sd_r = bf16_round_state_dict(sd)
cfg = cases.oracle_cfg(name)
slots, target = cases.make_inputs(name)
m = build_product(name); m.load_state_dict(sd, strict=False); m = m.to(torch.bfloat16).to(dev).eval()
ps = to_product_slots(slots, dev)
om.STORE_BF16 = True
R = lambda a, b: f"{rel_l2(a.float().cpu(), b):.3e}"
with torch.no_grad():
    embed, masks, pos, biases, _ = m.encoder.adaptor([s for s in ps if s.is_src])
    oe, omasks, opos, obiases = om.general_adaptor(sd_r, "encoder.adaptor", cfg, [s for s in slots if s.is_src], True)
    print("embed", R(embed, oe), " pos", R(pos, opos) if pos is not None else "-")
    layer = m.encoder.layers[li]
    p = f"encoder.layers.{li}"
    x = embed  # B x T x C fp32
    B, T, C = x.shape
    H, dh = cfg.heads, cfg.head_dim
    bias = biases[li] if biases is not None else None
    # dense bias tile of the product vs the oracle's bias
    if bias is not None:
        ab = bias.abs
        print("abs term (unscaled pq.pk)", R(ab[:, :, :T], (om._st(om.linear(opos, sd_r, "encoder.adaptor.pos_q_linear")).view(B, T, H, dh).transpose(1, 2)[0] @ om._st(om.linear(opos, sd_r, "encoder.adaptor.pos_k_linear")).view(B, T, H, dh).transpose(1, 2)[0].transpose(1, 2))))
    x1 = layer.self_attn_layer_norm(x)
    o_x1 = om.layer_norm(x.float().cpu(), sd_r, p + ".self_attn_layer_norm", st=True)
    print("pre-LN", R(x1, o_x1))
    a = layer.self_attn
    qkv = ops.linear(x1, a._cat(("q_proj", "k_proj", "v_proj"), "weight"), a._cat(("q_proj", "k_proj", "v_proj"), "bias"))
    xin = x1.float().cpu()
    oq, ok, ov = (om.linear(xin, sd_r, f"{p}.self_attn.{n}_proj") for n in "qkv")
    print("q", R(qkv[..., :C], oq), " k", R(qkv[..., C:2 * C], ok), " v", R(qkv[..., 2 * C:], ov))
    fast = bias is None
    scale = float(dh) ** -0.5 if fast else float(dh * cfg.attn_scale_factor) ** -0.5
    ctx = ops.attention(qkv, None, H, scale, bias, masks, False)
    # oracle attention on the product's q, k, v
    q_, k_, v_ = (qkv[..., i * C:(i + 1) * C].float().cpu().view(B, T, H, dh).permute(0, 2, 1, 3) for i in range(3))
    w = (q_ * scale) @ k_.transpose(2, 3)
    if obiases is not None:
        w = w + om._st_bias(obiases[li])
    if bool(omasks.any()):
        w = w.masked_fill(omasks[:, None, None, :], float("-inf"))
    x2l = w * 1.4426950408889634
    mm = torch.ceil(x2l.amax(-1, keepdim=True)); mm = torch.where(torch.isinf(mm), torch.zeros_like(mm), mm)
    pu = torch.exp2(x2l - mm)
    oc = om._st((om._st(pu) @ v_) / pu.sum(-1, keepdim=True)).permute(0, 2, 1, 3).reshape(B, T, C)
    print("attention ctx", R(ctx, oc))
    if bias is not None:
        # product's own bias tile (fp16, log2 domain) vs oracle bias
        pass
    w_out = a.out_proj.weight if fast or a.c_attn is None else ops.scale_cols(a.out_proj.weight, a.c_attn, dh)
    out = ops.linear(ctx, w_out, a.out_proj.bias)
    cin = ctx.float().cpu()
    if fast or a.c_attn is None:
        o_out = om.linear(cin, sd_r, p + ".self_attn.out_proj")
    else:
        we = om._st(sd_r[p + ".self_attn.out_proj.weight"] * sd_r[p + ".self_attn.c_attn"].repeat_interleave(dh).unsqueeze(0))
        print("scaled out_proj weight", R(w_out, we))
        o_out = om._st(F.linear(cin, we, sd_r[p + ".self_attn.out_proj.bias"]))
    print("out_proj", R(out, o_out))
    xn, x2 = ops.ln_res_ln(out, x, layer.attn_ln.weight, layer.attn_ln.bias, layer.final_layer_norm.weight, layer.final_layer_norm.bias, 1e-5)
    o_xn = x.float().cpu() + om.layer_norm(out.float().cpu(), sd_r, p + ".attn_ln")
    print("residual after attention", R(xn, o_xn), " final_layer_norm", R(x2, om.layer_norm(xn.float().cpu(), sd_r, p + ".final_layer_norm", st=True)))
    h = ops.linear(x2, layer.fc1.weight, layer.fc1.bias)
    print("fc1", R(h, om.linear(x2.float().cpu(), sd_r, p + ".fc1")))
    h2 = ops.layer_norm(h, layer.ffn_layernorm.weight, layer.ffn_layernorm.bias, 1e-5, gelu=True)
    print("gelu + ffn_layernorm", R(h2, om.layer_norm(om.gelu(h.float().cpu()), sd_r, p + ".ffn_layernorm", st=True)))
    y = ops.linear(h2, layer.fc2.weight, layer.fc2.bias)
    print("fc2", R(y, om.linear(h2.float().cpu(), sd_r, p + ".fc2")))
om.STORE_BF16 = False
