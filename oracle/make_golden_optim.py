"""TEST INFRASTRUCTURE ONLY (container only: needs /root/reference) -- golden vectors for the optimizer step.

Runs the reference's own `Adam` (engine/optim/adam.py, a torch.optim.Optimizer) and `clip_grad_norm_`
(module/utils.py:342-384), imported UNMODIFIED from /root/reference through the stub packages of oracle/ref_shim.py,
on the seeded case of oracle_optim.make_case(), with the deferred multiply factor of the bf16 wrapper
(fp16_optimizer.py:152-204) written out in this script.  Output: tests/golden/optim_adam.pt.

    python -m oracle.make_golden_optim
"""
import importlib
import os
import sys
import types

import torch

from . import oracle_optim, ref_shim


def _reference():
    ref_shim.install()
    eng = types.ModuleType("ofasys.engine")
    eng.__path__ = []  # no heavy __init__ (it imports criteria, EMA, checkpoint utils ...)
    sys.modules["ofasys.engine"] = eng
    opt = types.ModuleType("ofasys.engine.optim")
    opt.__path__ = [os.path.join(ref_shim.REF_PKG, "engine", "optim")]
    opt.FairseqOptimizer = object
    opt.register_optimizer = lambda *a, **k: (lambda cls: cls)
    sys.modules["ofasys.engine.optim"] = opt
    adam = importlib.import_module("ofasys.engine.optim.adam")
    utils = importlib.import_module("ofasys.module.utils")
    return adam.Adam, utils.clip_grad_norm_


def main():
    Adam, clip_grad_norm_ = _reference()
    params, steps, hyper, scales = oracle_optim.make_case()
    masters = [torch.nn.Parameter(p.float()) for p in params]  # fp32 copies (fp16_optimizer.py build_fp32_params)
    opt = Adam(masters, lr=hyper["lr"], betas=hyper["betas"], eps=hyper["eps"], weight_decay=hyper["weight_decay"])
    out = {"norms": [], "masters": [], "exp_avg": [], "exp_avg_sq": [], "params_bf16": []}
    for gs, c in zip(steps, scales):
        for p32, g in zip(masters, gs):
            p32.grad = torch.zeros_like(p32) if g is None else g.float()  # fp16_optimizer.py:104-131
        factor = float(c)  # multiply_grads :170-172
        grad_norm = factor * clip_grad_norm_(masters, 0)  # :178
        clip_coef = (hyper["max_norm"] / (grad_norm + 1e-6)).clamp_(max=1)  # :185-187
        factor = factor * clip_coef
        for p32 in masters:
            p32.grad.mul_(factor)  # :152-168
        opt.step()
        out["norms"].append(grad_norm.clone())
        out["masters"].append([p.detach().clone() for p in masters])
        out["exp_avg"].append([opt.state[p]["exp_avg"].clone() for p in masters])
        out["exp_avg_sq"].append([opt.state[p]["exp_avg_sq"].clone() for p in masters])
        out["params_bf16"].append([p.detach().to(torch.bfloat16) for p in masters])  # :134-150
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "optim_adam.pt")
    # keep the fixture small: full tensors for the small parameters, checksums for the two large ones
    small = [i for i, p in enumerate(params) if p.numel() <= 5000]
    fx = {"small": small, "norms": out["norms"]}
    for k in ("masters", "exp_avg", "exp_avg_sq", "params_bf16"):
        fx[k] = [{i: t[i] for i in small} for t in out[k]]
        fx[k + "_sum"] = [[t[i].double().sum() for i in range(len(params))] for t in out[k]]
        fx[k + "_abs"] = [[t[i].double().abs().sum() for i in range(len(params))] for t in out[k]]
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path), "bytes; norms", [float(n) for n in out["norms"]])


if __name__ == "__main__":
    main()
