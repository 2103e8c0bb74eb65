"""2D-CTC head with the reference's surface (decoders/ctc_decoder2d.py:7-53): mask branch (softmax over H) x
classify branch (softmax over C) -> log(max(., tiny)) -> (T,H,N,C) -> ops.ctc_loss_2d / target_length.
State-dict keys: saved_tiny, pred_mask.{1,2}.*, pred_classify.{1,2}.* (SURVEY.md App. C).  The two 3x3->1x1
conv branches run through ATen in this revision; the loss is megreader_b200's CUDA op."""
import torch
import torch.nn as nn

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

    def forward(self, feature, targets=None, lengths=None, train=False, masks=None, segs=None):
        tiny = self.saved_tiny
        if isinstance(feature, tuple):
            feature = feature[-1]
        mask = self.pred_mask(feature)
        classify = nn.functional.softmax(self.pred_classify(feature), dim=1)
        if self.training:
            pred = torch.log(torch.max(mask * classify, tiny))          # N, C, H, W
            pred = pred.permute(3, 2, 0, 1).contiguous()                # W, H, N, C
            input_lengths = torch.full((feature.size(0),), pred.shape[0], dtype=torch.long, device=pred.device)
            loss = self.ctc_loss(pred, targets.long(), input_lengths, lengths.long()) / lengths.float()
            return loss, pred
        return classify, mask
