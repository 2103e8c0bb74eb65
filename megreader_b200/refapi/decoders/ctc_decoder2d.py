"""2D-CTC head with the reference's surface (decoders/ctc_decoder2d.py:7-53): mask branch (softmax over H) x
classify branch (softmax over C) -> log(max(., tiny)) -> (T,H,N,C) -> ops.ctc_loss_2d / target_length.
State-dict keys: saved_tiny, pred_mask.{1,2}.*, pred_classify.{1,2}.* (SURVEY.md App. C).  The two 3x3->1x1
conv branches run through ATen; everything after them in the training branch (both softmaxes, product, max, log,
permuting copy, 2D-CTC loss and their backward) is ONE autograd node on megreader_b200's CUDA kernels
(megreader_b200/ctc2d_head.py, csrc/ctc2d_head.cu + csrc/ctc2d.cu): d(log_probs) never exists in HBM."""
import torch
import torch.nn as nn

from megreader_b200 import ctc2d_head
from megreader_b200.charset import default_charset


class CTCDecoder2D(nn.Module):
    def __init__(self, in_channels, charset=None, inner_channels=256, stride=1, blank=0, **kwargs):
        super().__init__()
        charset = charset if charset is not None else default_charset()
        self.charset = charset
        from ops import ctc_loss_2d          # same late import as the reference (:12)
        self.ctc_loss = ctc_loss_2d
        self.inner_channels = inner_channels
        self.pred_mask = nn.Sequential(
            nn.AvgPool2d(kernel_size=(stride, stride), stride=(stride, stride)),
            nn.Conv2d(in_channels, inner_channels, kernel_size=3, padding=1),
            nn.Conv2d(inner_channels, 1, kernel_size=1),
            nn.Softmax(dim=2))
        self.pred_classify = nn.Sequential(
            nn.AvgPool2d(kernel_size=(stride, stride), stride=(stride, stride)),
            nn.Conv2d(in_channels, inner_channels, kernel_size=3, padding=1),
            nn.Conv2d(inner_channels, len(charset), kernel_size=1))
        self.blank = blank
        self.tiny = torch.tensor(torch.finfo().tiny, requires_grad=False)
        self.register_buffer('saved_tiny', self.tiny)
        self._tiny_key, self._tiny_val = None, float(self.tiny)

    def _tiny(self):
        """Host copy of the `saved_tiny` buffer (the reference clamps with the BUFFER, so a checkpoint can change it);
        re-read only when the buffer object or its in-place version counter changed: no device read-back per step."""
        key = (id(self.saved_tiny), self.saved_tiny._version)
        if key != self._tiny_key:
            self._tiny_key, self._tiny_val = key, float(self.saved_tiny)
        return self._tiny_val

    def forward(self, feature, targets=None, lengths=None, train=False, masks=None, segs=None):
        if isinstance(feature, tuple):
            feature = feature[-1]
        mask_logits = self.pred_mask[2](self.pred_mask[1](self.pred_mask[0](feature)))    # before Softmax(dim=2)
        cls_logits = self.pred_classify(feature)
        if self.training:
            n, width = feature.size(0), cls_logits.size(3)
            input_lengths = torch.full((n,), width, dtype=torch.long, device=cls_logits.device)
            from ops import ctc_loss_2d
            if self.ctc_loss is ctc_loss_2d:
                nll, pred = ctc2d_head.head_loss(mask_logits, cls_logits, targets.long(), input_lengths, lengths.long(),
                                                 0, self._tiny())
            else:                                        # a caller swapped the loss: keep the epilogue, call theirs
                pred = ctc2d_head.head_log_probs(mask_logits, cls_logits, self._tiny())
                nll = self.ctc_loss(pred, targets.long(), input_lengths, lengths.long())
            return nll / lengths.float(), pred
        return nn.functional.softmax(cls_logits, dim=1), self.pred_mask[3](mask_logits)
