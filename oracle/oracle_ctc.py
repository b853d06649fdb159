"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CTC term of the ASR criterion.

Reference call site: ofasys/engine/criterion/speech_to_text_loss.py:339-379
    lprobs = log_softmax(logits.float(), -1)            # T x B x C, logits = F.linear(encoder_out, E[phone range]) (:215-224)
    F.ctc_loss(lprobs, targets_flat, input_lengths, target_lengths, blank=blank_idx, reduction="sum", zero_infinity=...)
The algorithm itself is PyTorch's (third-party, torch 2.11: aten/src/ATen/native/LossCTC.cpp -- Graves et al. 2006 alpha
recursion in log space).  This file restates it with differentiable torch ops (gradients by autograd, in float64 when
asked), so it also serves as the gradient oracle.  Pinned by tests/test_oracle_golden.py against F.ctc_loss itself and
against tests/golden/ctc.pt (F.ctc_loss outputs, oracle/make_golden_ctc.py)."""
import torch


def ctc_nll(lprobs, targets, input_lengths, target_lengths, blank=0):
    """lprobs [T, B, C] log-probabilities; targets [B, Lmax] left-aligned; returns nll [B] (inf when no alignment exists)."""
    T, B, C = lprobs.shape
    out = []
    # "log 0" is a huge negative finite number, so autograd never meets inf - inf; an nll above 1e29 means no alignment
    NEG = -1e30
    neg = torch.tensor(NEG, dtype=lprobs.dtype)
    inf = torch.tensor(float("inf"), dtype=lprobs.dtype)
    for b in range(B):
        Tb, L = int(input_lengths[b]), int(target_lengths[b])
        lab = targets[b, :L].long()
        ext = torch.full((2 * L + 1,), blank, dtype=torch.long)
        ext[1::2] = lab  # l': blanks at even positions
        S = ext.numel()
        if Tb == 0:
            out.append(torch.zeros((), dtype=lprobs.dtype) if L == 0 else inf)
            continue
        can_skip = torch.zeros(S, dtype=torch.bool)  # s-2 -> s allowed: l'_s is a label and differs from l'_{s-2}
        if S > 2:
            can_skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
        alpha = torch.where(torch.arange(S) < 2, lprobs[0, b, ext], neg.expand(S))
        for t in range(1, Tb):
            a1 = torch.cat([neg.view(1), alpha[:-1]])
            a2 = torch.cat([neg.view(1).expand(2), alpha[:-2]]) if S > 2 else neg.expand(S)
            a2 = torch.where(can_skip, a2, neg)
            alpha = torch.logsumexp(torch.stack([alpha, a1, a2]), dim=0).clamp_min(NEG) + lprobs[t, b, ext]
        nll = -torch.logsumexp(alpha[-2:] if S >= 2 else alpha[-1:], dim=0)
        out.append(nll if float(nll) < 1e29 else inf)
    return torch.stack(out)


def ctc_loss_sum(logits, targets, input_lengths, target_lengths, blank=0, zero_infinity=True):
    """speech_to_text_loss.py:339-379 on logits [T, B, C]: fp32 log-softmax + CTC, reduction "sum"."""
    lprobs = torch.log_softmax(logits.to(torch.float64 if logits.dtype == torch.float64 else torch.float32), dim=-1)
    nll = ctc_nll(lprobs, targets, input_lengths, target_lengths, blank)
    if zero_infinity:  # infinite losses and their gradients are zeroed: leave those utterances out of the graph
        keep = torch.isfinite(nll.detach())
        total = nll[keep].sum() if bool(keep.any()) else nll.new_zeros(())
        return total, torch.where(keep, nll.detach(), torch.zeros_like(nll.detach()))
    return nll.sum(), nll


def make_case(seed=5, T=37, B=4, C=29, Lmax=9, blank=1):
    """Seeded logits (bf16 values) and ragged labels; utterance 2 has repeated labels, utterance 3 is too short for its
    transcript (no alignment: infinite nll, zeroed by zero_infinity)."""
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(T, B, C, generator=g) * 1.5).to(torch.bfloat16).float()
    targets = torch.randint(2, C, (B, Lmax), generator=g)
    targets[2, :6] = torch.tensor([7, 7, 7, 3, 3, 9])
    in_len = torch.tensor([T, T - 5, 20, 4][:B], dtype=torch.long)
    tgt_len = torch.tensor([Lmax, 5, 6, 8][:B], dtype=torch.long)
    return logits, targets, in_len, tgt_len, blank
