"""Where does the CUDA path leave the oracle's storage model?  Per-layer residual-stream error of the encoder (and the final
logits) against the fp32 oracle and against the oracle under STORE_BF16, for one small case."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cases, oracle_model as om
from util import bf16_round_state_dict, build_product, load_golden, rel_l2, to_product_slots

name = sys.argv[1] if len(sys.argv) > 1 else "text_B"
dev = torch.device("cuda:0")
g = load_golden(name)
sd = cases.synth_state_dict(g["spec"], seed=0)
sd_r = bf16_round_state_dict(sd)
cfg = cases.oracle_cfg(name)
slots, target = cases.make_inputs(name)
m = build_product(name); m.load_state_dict(sd, strict=False); m = m.to(torch.bfloat16).to(dev).eval()
ps = to_product_slots(slots, dev)
with torch.no_grad():
    enc = m.encoder([s for s in ps if s.is_src], return_all_hiddens=True)
    logits, extra = m(ps, return_all_hiddens=True)

def oracle_states(store):
    om.STORE_BF16 = store
    try:
        with torch.no_grad():
            src = [s for s in slots if s.is_src]
            embed, masks, pos, biases = om.general_adaptor(sd_r, "encoder.adaptor", cfg, src, True)
            if bool(masks.any()):
                embed = embed * (1 - masks.unsqueeze(-1).type_as(embed))
            x = embed.transpose(0, 1)
            st = [x]
            for i in range(cfg.enc_layers):
                bias = biases[i].reshape(-1, x.size(0), x.size(0)) if biases is not None else None
                x = om.encoder_layer(sd_r, f"encoder.layers.{i}", cfg, x, masks if bool(masks.any()) else None, bias)
                st.append(x)
            xe = om.layer_norm(x, sd_r, "encoder.layer_norm", st=True)
            encd = {"encoder_out": xe, "encoder_padding_mask": masks, "position_embeddings": pos}
            tgt = [s for s in slots if not s.is_src]
            embed, dmasks, dpos, dbiases = om.general_adaptor(sd_r, "decoder.adaptor", cfg, tgt, False)
            B, T = embed.shape[:2]
            cross = None
            if cfg.mode == "A":
                sc = float(cfg.embed_dim / cfg.heads * cfg.attn_scale_factor) ** -0.5
                S = pos.shape[1]
                pq = om.linear(dpos, sd_r, "decoder.cross_pos_q_linear").view(B, T, cfg.heads, -1).transpose(1, 2) * sc
                pk = om.linear(pos, sd_r, "decoder.cross_pos_k_linear").view(B, S, cfg.heads, -1).transpose(1, 2)
                cross = torch.matmul(pq, pk.transpose(2, 3)).reshape(-1, T, S)
            y = embed.transpose(0, 1)
            dst = [y]
            future = torch.triu(torch.full((T, T), float("-inf")), 1)
            for i in range(cfg.dec_layers):
                bias = dbiases[i].reshape(-1, T, T) if dbiases is not None else None
                y = om.decoder_layer(sd_r, f"decoder.layers.{i}", cfg, y, xe, masks, future, dmasks, bias, cross)
                dst.append(y)
            lg, _ = om.model_forward(sd_r, cfg, slots)
        return st, lg, dst
    finally:
        om.STORE_BF16 = False

s32, l32, d32 = oracle_states(False)
s16, l16, d16 = oracle_states(True)
for i, t in enumerate(enc["encoder_states"]):
    t = t.float().cpu()
    print(f"enc state {i}: vs fp32 {rel_l2(t, s32[i]):.3e}   vs storage model {rel_l2(t, s16[i]):.3e}   (storage model vs fp32 {rel_l2(s16[i], s32[i]):.3e})")
for i, t in enumerate(extra["inner_states"]):
    t = t.float().cpu()
    print(f"dec state {i}: vs fp32 {rel_l2(t, d32[i]):.3e}   vs storage model {rel_l2(t, d16[i]):.3e}   (storage model vs fp32 {rel_l2(d16[i], d32[i]):.3e})")
lg = logits.float().cpu()
print(f"logits: vs fp32 {rel_l2(lg, l32):.3e}   vs storage model {rel_l2(lg, l16):.3e}   (storage model vs fp32 {rel_l2(l16, l32):.3e})")
