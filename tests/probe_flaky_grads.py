"""Repeat plain vs split backward of text_A and report every early parameter whose gradient is not bit-identical."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cases
from util import build_product, load_golden, to_product_slots
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "text_A"
g = load_golden(name); sd = cases.synth_state_dict(g["spec"], seed=0)
m = build_product(name); m.load_state_dict(sd, strict=False); m = m.to(torch.bfloat16).to(dev).train()
slots, target = cases.make_inputs(name); pslots = to_product_slots(slots, dev); tgt = target.to(dev)
names = {id(p): k for k, p in m.named_parameters()}
bad = {}
junk = []
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    junk.append(torch.empty(1000 + 37 * it, device=dev))  # perturb the allocator state between repetitions
    if it % 3 == 0: junk.clear()
    m.zero_grad(set_to_none=True)
    m.forward_loss(pslots, tgt).backward()
    ref = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    m.forward_loss(pslots, tgt).backward()
    for k, p in m.named_parameters():
        if p.grad is not None and not torch.equal(p.grad, ref[k]):
            d = (p.grad.float() - ref[k].float()).abs()
            bad.setdefault(("plain-vs-plain", k), []).append((int((d > 0).sum()), float(d.max()), float(ref[k].float().abs().max())))
    m.zero_grad(set_to_none=True)
    loss2, early, finish = m.forward_backward_split(pslots, tgt)
    for p in early:
        k = names[id(p)]
        if k in ref and not torch.equal(p.grad, ref[k]):
            d = (p.grad.float() - ref[k].float()).abs()
            bad.setdefault(("split-vs-plain", k), []).append((int((d > 0).sum()), float(d.max()), float(ref[k].float().abs().max())))
    finish()
torch.cuda.synchronize()
print("PDL", os.environ.get("OFAB_PDL", "1"), "mismatching (kind, param): count of repetitions, first records")
for k, v in sorted(bad.items()):
    print(k, len(v), v[:3])
print("done", len(bad))
