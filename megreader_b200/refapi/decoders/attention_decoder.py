"""Attention recogniser head with the reference's surface (decoders/attention_decoder.py:10-231).

Same modules, parameter names and shapes (state-dict compatible):
  encode.{0,1,3,4,6,7,9}.{0,1}   conv+BN(+ReLU) stack with pools (2,2),(2,1),(2,1) and a closing (2,3)/(2,1) conv
  onehot_embedding_{x,y}          identity-initialised position embeddings (max_size / height wide)
  decoder.{embedding,word_linear,attn.attn,attn.v,rnn,out}   Bahdanau cell: Linear(2H+E -> H), v, GRUCell(2H+E -> H)

What is done differently (same numbers up to fp32 re-association):
  * the additive-attention energy is Linear([h ; enc_l]) for every position l; the encoder half of that product
    does not depend on the step, so it is computed ONCE per call ((N,L,E)x(E,H)) and each step only adds the
    (N,H) hidden half — the reference rebuilds the (N,L,2H+E) concatenation and redoes the full product on each
    of its 32 steps (attention_decoder.py:152-169);
  * on CUDA tensors the recurrent loop -- training (with its backward through time) and greedy decoding -- runs as persistent
    cooperative kernels (megreader_b200/attn.py, csrc/attn_decode.cu); the framework composition below is what they are tested against;
  * the eval loop issues all steps without a host round-trip and applies the reference's early exit
    ("stop once every sample emitted blank", attention_decoder.py:129-130) afterwards on the device: columns
    after the first all-blank step are blank, which is exactly what the break leaves behind.
Random draws during training (teacher-forcing coin, step dropout) consume numpy / torch CPU generators in the
reference's order (attention_decoder.py:106-114), so seeded runs line up; they are made before the loop (draw_feedback) and
applied on the device (torch.where), so the loop has no host-dependent control flow and can be replayed from a CUDA graph.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from megreader_b200.charset import default_charset


class Attn(nn.Module):
    def __init__(self, method, hidden_dims, embed_size):
        super().__init__()
        self.method = method
        self.hidden_dims = hidden_dims
        self.embed_size = embed_size
        self.attn = nn.Linear(2 * hidden_dims + embed_size, hidden_dims)
        self.v = nn.Parameter(torch.empty(hidden_dims).normal_(mean=0, std=1.0 / math.sqrt(hidden_dims)))

    def project_encoder(self, encoder_outputs):
        """Step-invariant half of the energy: (L,N,H+E) -> (N,L,H), bias included."""
        w_enc = self.attn.weight[:, self.hidden_dims:]
        return torch.matmul(encoder_outputs.transpose(0, 1), w_enc.t()) + self.attn.bias

    def forward(self, hidden, encoder_outputs, projected=None):
        """hidden (N,H) [or (1,N,H)], encoder_outputs (L,N,H+E) -> attention weights (N,1,L)."""
        if projected is None:
            projected = self.project_encoder(encoder_outputs)
        hidden = hidden.reshape(-1, self.hidden_dims)
        from_hidden = F.linear(hidden, self.attn.weight[:, :self.hidden_dims])        # (N,H)
        energy = torch.tanh(projected + from_hidden.unsqueeze(1))                        # (N,L,H)
        return F.softmax(torch.matmul(energy, self.v), dim=1).unsqueeze(1)

    def score(self, hidden, encoder_outputs):
        """Reference-shaped entry (attention_decoder.py:160-171): hidden (N,L,H), encoder_outputs (N,L,H+E) -> (N,L)."""
        energy = torch.tanh(self.attn(torch.cat([hidden, encoder_outputs], 2)))
        return torch.matmul(energy, self.v)


class AttentionRNNCell(nn.Module):
    def __init__(self, hidden_dims, embedded_dims, nr_classes, n_layers=1, dropout_p=0, bidirectional=False):
        super().__init__()
        self.hidden_dims = hidden_dims
        self.embedded_dims = embedded_dims
        self.nr_classes = nr_classes
        self.n_layers = n_layers
        self.dropout_p = dropout_p
        self.embedding = nn.Embedding(nr_classes, nr_classes)
        self.embedding.weight.data = torch.eye(nr_classes)
        self.dropout = nn.Dropout(dropout_p)
        self.word_linear = nn.Linear(nr_classes, hidden_dims)
        self.attn = Attn('concat', hidden_dims, embedded_dims)
        self.rnn = nn.GRUCell(2 * hidden_dims + embedded_dims, hidden_dims)
        self.out = nn.Linear(hidden_dims, nr_classes)

    def forward(self, word_input, last_hidden, encoder_outputs, train=True, projected=None, encoder_bt=None):
        """One decoding step: word_input (N,), last_hidden (N,H), encoder_outputs (L,N,H+E)
        -> (log-probs or probs (N,V), hidden (N,H), attention (N,1,L))."""
        n = word_input.size(0)
        word = self.word_linear(self.embedding(word_input.to(last_hidden.device).long()))     # (N,H)
        weights = self.attn(last_hidden, encoder_outputs, projected)
        if encoder_bt is None:
            encoder_bt = encoder_outputs.transpose(0, 1)
        context = torch.bmm(weights, encoder_bt).squeeze(1)                                    # (N,H+E)
        hidden = self.rnn(torch.cat([word, context], 1), last_hidden.view(n, -1))
        logits = self.out(hidden)
        return (F.log_softmax(logits, dim=1) if train else F.softmax(logits, dim=1)), hidden, weights


class AttentionDecoder(nn.Module):
    def __init__(self, in_channels, charset=None, inner_channels=512, max_size=32, height=1, gt_as_output=None,
                 step_dropout=0, **kwargs):
        super().__init__()
        self.inner_channels = inner_channels
        self.encode = self._init_encoder(in_channels)
        self.max_size = max_size
        self.charset = charset if charset is not None else default_charset()
        self.height = height
        self.decoder = AttentionRNNCell(inner_channels, max_size + height, len(self.charset))
        self.step_dropout = step_dropout
        self.onehot_embedding_x = nn.Embedding(max_size, max_size)
        self.onehot_embedding_x.weight.data = torch.eye(max_size)
        self.onehot_embedding_y = nn.Embedding(height, height)
        self.onehot_embedding_y.weight.data = torch.eye(height)
        self.gt_as_output = gt_as_output
        self.feedback_static = None       # (coin, swap, noise) device tensors of a CUDA-graph caller, see forward()
        # CUDA tensors: the recurrent loop (training and greedy decoding) runs on csrc/attn_decode.cu.  False = the framework
        # composition (what the kernels are tested against; also the way out for hidden sizes whose weight slices do not fit in
        # shared memory, inner_channels > ~640, which the kernels refuse with MR_ERR_UNSUPPORTED)
        self.loop_kernels = True
        self.loss_function = nn.NLLLoss(reduction='none')

    def conv_bn_relu(self, input_channels, output_channels, kernel_size=3, stride=1, padding=1):
        return nn.Sequential(nn.Conv2d(input_channels, output_channels, kernel_size, stride, padding),
                             nn.BatchNorm2d(output_channels), nn.ReLU(inplace=True))

    def _init_encoder(self, in_channels, stride=(2, 1), padding=(0, 1)):
        c = self.inner_channels
        cbr = self.conv_bn_relu
        return nn.Sequential(cbr(in_channels, c), cbr(c, c), nn.MaxPool2d((2, 2), (2, 2), (0, 0)),
                             cbr(c, c), cbr(c, c), nn.MaxPool2d(stride, stride, (0, 0)),
                             cbr(c, c), cbr(c, c), nn.MaxPool2d(stride, stride, (0, 0)),
                             cbr(c, c, kernel_size=(2, 3), stride=stride, padding=padding))

    def _get_gt_as_output(self):
        if self.gt_as_output is not None:
            return self.gt_as_output
        return np.random.rand() < 0.5

    def draw_feedback(self, batch):
        """The random draws of one training forward, made on the host in the reference's order (per step: the numpy teacher-forcing
        coin, then torch.rand and torch.randint of the step dropout; attention_decoder.py:106-114):
        coin (max_size,) bool, swap (max_size, batch) int64 in {0, 1}, noise (max_size, batch) int64 class indices."""
        vocab = len(self.charset)
        coin, swap, noise = [], [], []
        for _ in range(self.max_size):
            coin.append(bool(self._get_gt_as_output()))
            swap.append((torch.rand(batch) < self.step_dropout).long())
            noise.append(torch.randint(high=vocab, size=(batch,)))
        return torch.tensor(coin, dtype=torch.bool), torch.stack(swap), torch.stack(noise)

    def _positions(self, batch, device):
        """(N, height+max_size, height, max_size): one-hot row then column coordinates of every cell."""
        ys = torch.arange(self.height, device=device).view(-1, 1).expand(self.height, self.max_size)
        xs = torch.arange(self.max_size, device=device).view(1, -1).expand(self.height, self.max_size)
        ey = self.onehot_embedding_y(ys).permute(2, 0, 1)
        ex = self.onehot_embedding_x(xs).permute(2, 0, 1)
        return torch.cat([ey, ex], 0).unsqueeze(0).expand(batch, -1, -1, -1)

    def _decode_cuda(self, memory_bt, projected, want_prob=False):
        """The eval loop on the GPU: one persistent kernel for all max_size steps (csrc/attn_decode.cu).  memory_bt (N,L,H+E),
        projected (N,L,H) -> pred (N,max_size) int32 [, per-step softmax (N,max_size,V)]."""
        from megreader_b200 import _lib
        cell = self.decoder
        n, L, _ = memory_bt.shape
        H, E, V, S = self.inner_channels, self.max_size + self.height, len(self.charset), self.max_size
        dev = memory_bt.device
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731
        memory_bt, projected = f32(memory_bt), f32(projected)
        wa = f32(cell.attn.attn.weight)                                   # (H, 2H+E); the kernel reads [:, :H] through the row stride
        wordtab = f32(cell.word_linear(cell.embedding.weight))           # (V, H): row w = word_linear(embedding(w))
        params = [f32(t) for t in (cell.attn.v, cell.rnn.weight_ih, cell.rnn.bias_ih, cell.rnn.weight_hh, cell.rnn.bias_hh,
                                   cell.out.weight, cell.out.bias)]
        pred = torch.empty(n, S, dtype=torch.int32, device=dev)
        prob = torch.empty(n, S, V, dtype=torch.float32, device=dev) if want_prob else None
        lib = _lib.lib()
        ws_bytes = int(lib.mr_attn_decode_workspace_bytes(n, H, E))
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.mr_attn_decode_f32(
                projected.data_ptr(), memory_bt.data_ptr(), wa.data_ptr(), wa.stride(0), params[0].data_ptr(), wordtab.data_ptr(),
                params[1].data_ptr(), params[2].data_ptr(), params[3].data_ptr(), params[4].data_ptr(), params[5].data_ptr(),
                params[6].data_ptr(), pred.data_ptr(), prob.data_ptr() if prob is not None else None, ws.data_ptr(), ws_bytes,
                n, L, H, E, V, S, int(self.charset.blank), stream), "attn_decode")
            if not torch.cuda.is_current_stream_capturing():
                import ctypes
                status = ctypes.c_int(0)
                _lib.check(lib.mr_attn_decode_status(ws.data_ptr(), n, H, E, stream, ctypes.byref(status)), "attn_decode_status")
                if status.value:
                    raise RuntimeError("megreader_b200 attn_decode: grid barrier timed out (error word %d)" % status.value)
        return (pred, prob) if want_prob else pred

    def forward(self, feature, targets=None, lengths=None, train=False):
        device = feature.device
        n = feature.shape[0]
        # the conv encoder may run in bf16 (megreader_b200.conv_engine); the recurrent part is fp32 like its parameters
        grid = torch.cat([self.encode(feature).float(), self._positions(n, device)], dim=1)
        memory = grid.reshape(n, grid.shape[1], -1).permute(2, 0, 1)            # (L,N,H+E), L = height*max_size
        memory_bt = memory.transpose(0, 1)
        projected = self.decoder.attn.project_encoder(memory)
        blank = self.charset.blank
        hidden = grid.new_zeros(n, self.inner_channels)
        word = torch.full((n,), blank, dtype=torch.long, device=device)
        vocab = len(self.charset)

        if self.training:
            targets = targets.long()
            lengths = lengths.to(device)
            loss = None
            attention = []
            # the step's random draws (teacher-forcing coin, step dropout) as device tensors: either handed in by a caller that
            # replays this forward from a CUDA graph (`feedback_static`, refreshed from draw_feedback() before every replay) or
            # drawn here, on the host, in the reference's order (attention_decoder.py:106-114)
            coin, swap, noise = self.feedback_static if self.feedback_static is not None else \
                tuple(t.to(device) for t in self.draw_feedback(n))
            if feature.is_cuda and self.loop_kernels:
                # the whole loop (and its backward through time) on the persistent kernels of csrc/attn_decode.cu
                from megreader_b200 import attn as attn_kernels
                loss, maps = attn_kernels.attention_loop_loss(projected, memory_bt, self.decoder, targets, lengths,
                                                              (coin, swap, noise), blank)
                return loss, maps.view(n, -1, self.height, self.max_size)
            for t in range(self.max_size):
                logp, hidden, weights = self.decoder(word, hidden, memory, True, projected, memory_bt)
                step = self.loss_function(logp, targets[:, t]) * (t <= lengths).float()
                loss = step if loss is None else loss + step
                attention.append(weights)
                word = torch.where(coin[t], targets[:, t], logp.argmax(dim=1).detach())
                # step dropout: a random class replaces the fed-back symbol with probability step_dropout
                word = word * (1 - swap[t]) + noise[t] * swap[t]
            return loss, torch.cat(attention, 1).view(n, -1, self.height, self.max_size)

        if feature.is_cuda and self.loop_kernels:
            pred = self._decode_cuda(memory_bt, projected)
            finished = (pred == blank).all(dim=0).long().cummax(0).values.bool()
            return pred.masked_fill(finished.unsqueeze(0), blank)
        pred = self._decode_aten(memory, memory_bt, projected)
        finished = (pred == blank).all(dim=0).long().cummax(0).values.bool()   # step t or an earlier one was all-blank
        return pred.masked_fill(finished.unsqueeze(0), blank).to(torch.int32)

    def _decode_aten(self, memory, memory_bt, projected, want_prob=False):
        """The eval loop as a framework composition (CPU tensors; the comparison arm of tests/test_attention_decode_gpu.py)."""
        n = memory_bt.shape[0]
        hidden = memory_bt.new_zeros(n, self.inner_channels)
        word = torch.full((n,), self.charset.blank, dtype=torch.long, device=memory_bt.device)
        steps, probs = [], []
        for t in range(self.max_size):
            prob, hidden, _ = self.decoder(word, hidden, memory, False, projected, memory_bt)
            word = prob.argmax(dim=1)
            steps.append(word)
            probs.append(prob)
        pred = torch.stack(steps, 1)                                             # (N, max_size)
        return (pred, torch.stack(probs, 1)) if want_prob else pred
