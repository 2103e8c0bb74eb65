"""CRNN decoder with the reference's surface (decoders/crnn.py:8-104): BidirectionalLSTM x2 + Linear, then
train: log_softmax -> CTC (mean, zero_infinity) returning (loss, log_probs float64); eval: softmax -> (N,C,1,T).
State-dict keys: rnn.{0,1}.rnn.{weight_ih_l0,...,*_reverse}, rnn.{0,1}.embedding.{weight,bias} (App. C)."""
import torch
import torch.nn as nn

from megreader_b200 import crnn_engine
from megreader_b200.charset import default_charset


class BidirectionalLSTM(nn.Module):
    def __init__(self, nIn, nHidden, nOut):
        super().__init__()
        self.rnn = nn.LSTM(nIn, nHidden, bidirectional=True)
        self.embedding = nn.Linear(nHidden * 2, nOut)

    def forward(self, input):
        return crnn_engine.bilstm_forward(self, input)


class CRNNDecoder(nn.Module):
    def __init__(self, charset=None, inner_channels=256, in_channels=256, need_reduce=False, reduce_func=None,
                 loss_func='pytorch'):
        super().__init__()
        charset = charset if charset is not None else default_charset()
        self.charset = charset
        rnn_input = inner_channels if need_reduce else in_channels
        self.rnn = nn.Sequential(BidirectionalLSTM(rnn_input, inner_channels, inner_channels),
                                 BidirectionalLSTM(inner_channels, inner_channels, len(charset)))
        self.inner_channels = inner_channels
        if need_reduce:
            if reduce_func == 'conv':
                self.fpn2rnn = self._init_conv(in_channels)
            elif reduce_func == 'pooling':
                self.fpn2rnn = nn.AdaptiveMaxPool2d((1, None))
        # 'pytorch' -> nn.CTCLoss(zero_infinity=True) semantics (mean over batch of nll/len);
        # anything else -> the reference's python CTCLoss semantics (per-sample nll/len, decoders/ctc_loss.py:118-121)
        self.loss_func = loss_func

    def conv_bn_relu(self, cin, cout, kernel_size=3, stride=1, padding=1):
        return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=kernel_size, stride=stride, padding=padding),
                             nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def _init_conv(self, in_channels, stride=(2, 1), padding=(0, 1)):
        c = self.inner_channels
        return nn.Sequential(self.conv_bn_relu(in_channels, c), nn.MaxPool2d((2, 2), (2, 2), (0, 0)),
                             self.conv_bn_relu(c, c), nn.MaxPool2d(stride, stride, (0, 0)),
                             self.conv_bn_relu(c, c), nn.MaxPool2d(stride, stride, (0, 0)))

    def forward(self, feature, targets=None, lengths=None, train=False):
        return crnn_engine.decoder_forward(self, feature, targets, lengths, train)
