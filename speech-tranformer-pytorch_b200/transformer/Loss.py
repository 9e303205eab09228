"""B200 drop-in for the reference's transformer/Loss.py (LabelSmoothingLoss :6-39, CrossEntropyLoss :42-73).

Unlike the reference, which builds its temporaries with `torch.zeros(...)` on the CPU (Loss.py:61-65)
and therefore only runs on CPU tensors, this runs on the device the logits live on.  Semantics kept,
including the quirks: the `one_hot` column `ignore_index` is zeroed only when ignore_index == 0
(`if not ignore_index`, Loss.py:20), the mean divides by ALL rows (Loss.py:69), and a class `weight`
vector is required (Loss.py:59 dereferences it)."""
import torch
import torch.nn as nn

from .. import functional as F

__all__ = ["LabelSmoothingLoss", "CrossEntropyLoss", "CTCLoss", "JointCTCAttentionLoss"]


class LabelSmoothingLoss(nn.Module):
    def __init__(self, label_smoothing, vocab_size, weight=None, size_average=True, ignore_index=-1):
        assert 0.0 <= label_smoothing <= 1.0                       # Loss.py:14
        self.padding_idx = ignore_index
        super(LabelSmoothingLoss, self).__init__()

        smoothing_value = label_smoothing / (vocab_size - 1)
        one_hot = torch.full((vocab_size,), smoothing_value)
        if not ignore_index:                                       # Loss.py:20 (sic)
            one_hot[self.padding_idx] = 0
        self.register_buffer('one_hot', one_hot.unsqueeze(0))

        self.confidence = 1.0 - label_smoothing
        self.criterion = CrossEntropyLoss(weight=weight, size_average=size_average)

    def forward(self, output, target):
        """
        output (FloatTensor): batch_size x n_classes
        target (LongTensor): batch_size
        """
        weight = self.criterion.weight
        if weight is None:                                         # the reference fails at Loss.py:59
            raise AttributeError("'NoneType' object has no attribute 'repeat'")
        if weight.device != output.device:
            weight = weight.to(output.device)
        one_hot = self.one_hot if self.one_hot.device == output.device else self.one_hot.to(output.device)
        return F.label_smoothing_ce(output, target, one_hot, weight, self.confidence, self.padding_idx,
                                    self.criterion.size_average)


class CrossEntropyLoss(nn.Module):
    def __init__(self, weight, size_average=True):
        super(CrossEntropyLoss, self).__init__()
        self.weight = weight
        self.size_average = size_average
        self.log_softmax = nn.LogSoftmax(dim=-1)

    def forward(self, inputs, target):
        assert inputs.dim() == 2                                   # Loss.py:51-52
        assert target.dim() == 2
        if self.weight is None:
            raise AttributeError("'NoneType' object has no attribute 'repeat'")
        weight = self.weight if self.weight.device == inputs.device else self.weight.to(inputs.device)
        return F.soft_target_ce(inputs, target, weight, self.size_average)



class CTCLoss(nn.Module):
    """CTC head for the joint CTC / attention objective of BASELINE.json configs[3].  The reference's
    train_attn_and_ctc.py is an empty file, so the interface follows torch.nn.CTCLoss with batch-first frame LOGITS
    (B, T, V) — the log-softmax is fused — and padded (B, L_max) targets."""

    def __init__(self, blank=0, reduction="mean"):
        super(CTCLoss, self).__init__()
        self.blank, self.reduction = blank, reduction

    def forward(self, logits, targets, input_lengths, target_lengths):
        return F.ctc_loss(logits, targets, input_lengths, target_lengths, blank=self.blank, reduction=self.reduction)


class JointCTCAttentionLoss(nn.Module):
    """loss = ctc_weight * CTC(encoder-side logits) + (1 - ctc_weight) * attention criterion(decoder logits)."""

    def __init__(self, attention_criterion, ctc_weight=0.3, blank=0, ctc_reduction="mean"):
        super(JointCTCAttentionLoss, self).__init__()
        assert 0.0 <= ctc_weight <= 1.0
        self.att, self.ctc, self.ctc_weight = attention_criterion, CTCLoss(blank, ctc_reduction), ctc_weight

    def forward(self, dec_logits, dec_truth, ctc_logits, ctc_targets, input_lengths, target_lengths):
        att = self.att(dec_logits, dec_truth)
        ctc = self.ctc(ctc_logits, ctc_targets, input_lengths, target_lengths)
        return self.ctc_weight * ctc + (1.0 - self.ctc_weight) * att
