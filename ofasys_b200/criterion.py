"""Criteria at the boundary of the measured path, same call surface as the reference's
(`criterion(model, sample, update_num) -> (loss, sample_size, logging_output)`), executed by the CE / CTC kernels:

  LabelSmoothedCrossEntropyCriterion   engine/criterion/label_smoothed_cross_entropy.py:97-214 (label smoothing, ignore_prefix_size,
                                       constraint_range / sample["constraint_masks"], drop_worst_ratio / drop_worst_after)
  SpeechToTextLossCriterion            engine/criterion/speech_to_text_loss.py:133-380: ce_weight * label-smoothed CE +
                                       ctc_weight * CTC on F.linear(encoder_out, E[dict_start:dict_end]) (the tied embedding rows
                                       of the phone range as the CTC head)

`sample` is the reference's dict: {"net_input": {"slots": [...]}, "target": int64 [B, T], "ntokens": int, "nsentences": int,
optional "constraint_masks": bool [B, T, V], ASR: "encoder_target": int64 [B, L], optional "target_lengths"}.
"""
import torch

from . import ops


class LabelSmoothedCrossEntropyCriterion:
    def __init__(self, label_smoothing=0.0, ignore_prefix_size=0, drop_worst_ratio=0.0, drop_worst_after=0, constraint_range=None,
                 sentence_avg=False, padding_idx=1, weight=1.0, report_accuracy=False):
        self.eps = float(label_smoothing)
        self.ignore_prefix_size = int(ignore_prefix_size)
        self.drop_worst_ratio, self.drop_worst_after = float(drop_worst_ratio), int(drop_worst_after)
        self.constraint_start = self.constraint_end = None
        if constraint_range is not None:  # "start,end" as in the reference's config, or a pair
            a, b = constraint_range.split(",") if isinstance(constraint_range, str) else constraint_range
            self.constraint_start, self.constraint_end = int(a), int(b)
        self.sentence_avg, self.padding_idx, self.weight, self.report_accuracy = sentence_avg, padding_idx, weight, report_accuracy

    def __call__(self, model, sample, update_num=0, reduce=True):
        return self.forward(model, sample, update_num, reduce)

    def forward(self, model, sample, update_num=0, reduce=True):
        net_output = model(**sample["net_input"])
        loss, nll_loss, ntokens = self.compute_loss(model, net_output, sample, update_num, reduce=reduce)
        sample_size = sample["target"].size(0) if self.sentence_avg else ntokens
        logging_output = {"loss": loss.detach(), "nll_loss": nll_loss.detach(), "ntokens": sample.get("ntokens", ntokens),
                          "nsentences": sample.get("nsentences", sample["target"].size(0)), "sample_size": sample_size}
        if self.report_accuracy:
            n_correct, total = self.compute_accuracy(model, net_output, sample)
            logging_output["n_correct"], logging_output["total"] = int(n_correct), int(total)
        return loss * self.weight, sample_size, logging_output

    def _constraints(self, sample):
        rng = None if self.constraint_start is None else (self.constraint_start, self.constraint_end)
        return sample.get("constraint_masks", None), rng

    def compute_loss(self, model, net_output, sample, update_num, reduce=True):
        """label_smoothed_cross_entropy.py:159-191: log-softmax over the allowed entries, padding rows dropped, optional
        drop-worst (keep the int(n (1 - ratio)) smallest row losses after `drop_worst_after` updates), sums."""
        logits = net_output[0]
        target = model.get_targets(sample, net_output) if hasattr(model, "get_targets") else sample["target"]
        cmask, rng = self._constraints(sample)
        if self.ignore_prefix_size > 0:
            logits = logits[:, self.ignore_prefix_size:, :]
            target = target[:, self.ignore_prefix_size:]
            if cmask is not None:
                cmask = cmask[:, self.ignore_prefix_size:, :]
        row_loss, row_nll = ops.cross_entropy_rows(logits, target, self.padding_idx, self.eps, cmask, rng)
        keep = target.reshape(-1) != self.padding_idx
        loss_rows, nll_rows = row_loss[keep], row_nll[keep]
        if self.drop_worst_ratio > 0 and update_num > self.drop_worst_after:
            loss_rows, idx = torch.topk(loss_rows, k=int(loss_rows.shape[0] * (1 - self.drop_worst_ratio)), largest=False)
            nll_rows = nll_rows[idx]
        return loss_rows.sum(), nll_rows.sum(), loss_rows.numel()

    def compute_accuracy(self, model, net_output, sample):
        logits = net_output[0].float()
        cmask, rng = self._constraints(sample)
        if rng is not None:
            logits = logits.clone()
            logits[..., 4:rng[0]] = float("-inf")
            logits[..., rng[1]:] = float("-inf")
        if cmask is not None:
            logits = logits.masked_fill(~cmask, float("-inf"))
        target = sample["target"]
        mask = target.ne(self.padding_idx)
        return (logits.argmax(-1).eq(target) & mask).sum(), mask.sum()


class SpeechToTextLossCriterion(LabelSmoothedCrossEntropyCriterion):
    """speech_to_text_loss.py:133-380: `ce_weight * CE + ctc_weight * CTC`; the CTC logits are the encoder output projected on
    the rows [dict_start, dict_end) of the tied embedding (the phone vocabulary), blank = `blank_idx` within that range."""

    def __init__(self, dict_start, dict_end, blank_idx=0, ce_weight=1.0, ctc_weight=0.0, zero_infinity=True, eos_idx=2, **kw):
        super().__init__(**kw)
        self.dict_start, self.dict_end, self.blank_idx = int(dict_start), int(dict_end), int(blank_idx)
        self.ce_weight, self.ctc_weight, self.zero_infinity, self.eos_idx = float(ce_weight), float(ctc_weight), bool(zero_infinity), eos_idx

    def forward(self, model, sample, update_num=0, reduce=True):
        logits, extra, encoder_out = model(**sample["net_input"], return_encoder_out=True)
        loss_ce = nll_ce = None
        ntokens = sample["ntokens"] if "ntokens" in sample else None
        if self.ce_weight > 0:
            loss_ce, nll_ce, n = self.compute_loss(model, (logits, extra), sample, update_num, reduce=reduce)
            ntokens = n if ntokens is None else ntokens
        loss_ctc = None
        if self.ctc_weight > 0:
            loss_ctc = self.compute_loss_ctc(model, encoder_out, sample)
        if loss_ce is not None and loss_ctc is not None:
            loss = self.ce_weight * loss_ce + self.ctc_weight * loss_ctc
        else:
            loss = loss_ce if loss_ce is not None else loss_ctc
        if ntokens is None:
            ntokens = int(sample["target_lengths"].sum())
        sample_size = sample["target"].size(0) if self.sentence_avg else ntokens
        log = {"loss": loss.detach(), "ce_loss": 0 if loss_ce is None else loss_ce.detach(), "ctc_loss": 0 if loss_ctc is None else loss_ctc.detach(),
               "nll_loss": 0 if nll_ce is None else nll_ce.detach(), "ntokens": ntokens, "nsentences": sample["target"].size(0), "sample_size": sample_size}
        return loss, sample_size, log

    def compute_loss_ctc(self, model, encoder_out, sample):
        """speech_to_text_loss.py:339-379: log-softmax of the phone-range projection of the encoder output, input lengths from the
        encoder padding mask, targets = encoder_target - dict_start without pad / eos; F.ctc_loss(reduction="sum")."""
        E = model.decoder.adaptor.embed_tokens.weight[self.dict_start:self.dict_end]
        enc = encoder_out["_encoder_out_bt"] if "_encoder_out_bt" in encoder_out else encoder_out["encoder_out"][0].transpose(0, 1)
        x = ops.linear(ops.to_bf16(enc), E, None)  # B x T x C
        kpm = encoder_out["encoder_padding_mask"][0] if encoder_out["encoder_padding_mask"] else None
        input_lengths = None if kpm is None else (~kpm.bool()).long().sum(-1)
        et = sample["encoder_target"]
        pad_mask = (et != self.padding_idx) & (et != self.eos_idx)
        targets = et - self.dict_start
        target_lengths = (sample["target_lengths"] - 1) if "target_lengths" in sample else pad_mask.sum(-1)
        # left-align the kept labels per utterance (the reference passes them flattened; F.ctc_loss accepts both forms)
        B, L = et.shape
        order = torch.argsort((~pad_mask).to(torch.int8), dim=1, stable=True)
        packed = torch.gather(targets, 1, order)
        loss, _ = ops.ctc_loss_sum(x, packed, target_lengths, input_lengths, blank=self.blank_idx, zero_infinity=self.zero_infinity)
        return loss
