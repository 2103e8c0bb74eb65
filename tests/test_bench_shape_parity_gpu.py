"""GPU parity of the path bench.py actually times, AT the bench shape (BASELINE.json cfg 2: N = 512 lines of 3x32x256, T = 65):

  (a) the engine in fp32 mode against the CPU oracle port (oracle/crnn_port.py == the unmodified reference modules, bit for bit:
      tests/test_oracle_crnn.py) -- loss, log-probs and feature map within the north_star tolerance (1e-4 relative), and the
      greedy label indices bit-exact;
  (b) the engine in bf16 mode (the 57 k lines/s number) against that fp32 result: loss delta, largest log-prob delta and
      the arg-max AGREEMENT RATE over the T x N columns, asserted against thresholds and printed;
  (c) the persistent whole-sequence BiLSTM kernels at T = 65, N = 512, H = 256 against torch's nn.LSTM in fp32 (not against
      the repo's own step-mode kernels).

bench.py recomputes (b) live in its JSON line ("parity")."""
import numpy as np
import pytest
import torch

from tests.weights import crnn_batch, fill_state_dict

pytestmark = pytest.mark.gpu
N, W, T, L_MAX = 512, 256, 65, 16


@pytest.fixture(scope="module")
def nets(cuda):
    import megreader_b200
    megreader_b200.install_reference_api()
    import backbones
    import decoders
    bb = fill_state_dict(backbones.crnn_backbone(), "bb.").to(cuda)
    dec = fill_state_dict(decoders.CRNNDecoder(in_channels=512, inner_channels=256), "dec.").to(cuda)
    return bb, dec


@pytest.fixture(scope="module")
def batch():
    return crnn_batch(0, N, W, L_MAX, T)


def _engine_step(cuda, nets, batch, dtype):
    from megreader_b200 import crnn_engine
    bb, dec = nets
    bb.train(); dec.train()
    for p in list(bb.parameters()) + list(dec.parameters()):
        p.grad = None
    state = {k: v.clone() for k, v in bb.state_dict().items()}          # BN running stats must not drift between runs
    x, labels, lengths = [torch.from_numpy(a).to(cuda) for a in batch]
    crnn_engine.set_compute_dtype(dtype)
    try:
        feat = bb(x)
        loss, lp = dec(feat, targets=labels, lengths=lengths, train=True)
        loss.mean().backward()
        torch.cuda.synchronize()
        if crnn_engine.LAST_LSTM_FLAGS is not None:
            assert int(crnn_engine.LAST_LSTM_FLAGS[-1]) == 0, "persistent LSTM: inter-CTA wait timed out"
    finally:
        crnn_engine.set_compute_dtype(torch.float32)
    grads = {n: p.grad.detach().double().norm().item() for n, p in list(bb.named_parameters()) + list(dec.named_parameters())}
    bb.load_state_dict(state)
    return float(loss.item()), lp.detach().float().cpu().numpy(), feat.detach().float().cpu().numpy(), grads


@pytest.fixture(scope="module")
def fp32_run(cuda, nets, batch):
    return _engine_step(cuda, nets, batch, torch.float32)


def test_fp32_engine_vs_cpu_oracle_at_bench_shape(fp32_run, batch):
    """(a) whole 512-line batch through the CPU port (BatchNorm uses batch statistics, so the full batch is the unit)."""
    from oracle import crnn_port
    x, labels, lengths = [torch.from_numpy(a) for a in batch]
    torch.set_num_threads(max(1, min(32, torch.get_num_threads())))
    pb = fill_state_dict(crnn_port.CRNNBackbonePort(), "bb.").train()
    pd = fill_state_dict(crnn_port.CRNNDecoderPort(), "dec.").train()
    feat = pb(x)
    loss, lp = pd(feat, labels, lengths, train=True)
    loss.mean().backward()
    g_loss, g_lp, g_feat, g_grads = fp32_run
    np.testing.assert_allclose(g_loss, loss.item(), rtol=1e-4)
    np.testing.assert_allclose(g_feat, feat.detach().numpy(), rtol=1e-4, atol=1e-4)
    ref_lp = lp.detach().float().numpy()
    np.testing.assert_allclose(g_lp, ref_lp, rtol=1e-4, atol=1e-4)
    # label indices (north_star: bit-exact): arg-max per column wherever the reference's own top-2 margin is above fp32 noise
    top2 = np.sort(ref_lp, axis=2)[:, :, -2:]
    decided = (top2[:, :, 1] - top2[:, :, 0]) > 2e-4
    assert decided.mean() > 0.99
    assert np.array_equal(g_lp.argmax(2)[decided], ref_lp.argmax(2)[decided])
    cpu_g = {("bb." + n): p.grad.double().norm().item() for n, p in pb.named_parameters()}
    cpu_g.update({("dec." + n): p.grad.double().norm().item() for n, p in pd.named_parameters()})
    for n, v in g_grads.items():
        key = ("bb." if n.startswith("cnn") else "dec.") + n
        if cpu_g[key] > 1e-5:                       # conv bias in front of BatchNorm: true gradient 0, value = noise
            np.testing.assert_allclose(v, cpu_g[key], rtol=2e-3, err_msg=n)


def test_bf16_engine_vs_fp32_at_bench_shape(cuda, nets, batch, fp32_run, capsys):
    """(b) the benchmarked mode.  Random-init weights make the class posteriors nearly flat (loss ~ ln 38 per column), the
    hardest case for arg-max agreement; trained models have margins orders of magnitude larger."""
    loss32, lp32, feat32, g32 = fp32_run
    loss16, lp16, feat16, g16 = _engine_step(cuda, nets, batch, torch.bfloat16)
    rel = abs(loss16 - loss32) / abs(loss32)
    dmax = float(np.abs(lp16 - lp32).max())
    agree = float((lp16.argmax(2) == lp32.argmax(2)).mean())
    top2 = np.sort(lp32, axis=2)[:, :, -2:]
    margin = top2[:, :, 1] - top2[:, :, 0]
    decided = margin > 4 * dmax                                       # columns whose fp32 margin exceeds bf16 noise
    agree_decided = float((lp16.argmax(2)[decided] == lp32.argmax(2)[decided]).mean()) if decided.any() else 1.0
    with capsys.disabled():
        print("\n[bench-shape parity] bf16 vs fp32 engine: loss rel delta %.3e, max |log-prob delta| %.3e, arg-max agreement "
              "%.4f (%.4f on the %.1f%% of columns whose fp32 top-2 margin > 4x that delta)"
              % (rel, dmax, agree, agree_decided, 100 * decided.mean()))
    assert rel < 1e-2
    assert dmax < 0.25
    assert agree > 0.90
    assert agree_decided == 1.0
    for n, v in g16.items():
        if g32[n] > 1e-5:
            assert abs(v - g32[n]) / g32[n] < 0.1, n


def test_persistent_lstm_vs_torch_lstm_fp32_at_bench_shape(cuda):
    """(c) lstm_seq_fwd/bwd_kernel at T = 65, N = 512, H = 256 (both CRNN layers' geometry) vs nn.LSTM + nn.Linear in fp32."""
    from megreader_b200 import crnn_engine
    from tests.test_nn_kernels_gpu import _bilstm_case, _bilstm_run
    for (I, O, seed) in [(512, 256, 41), (256, 38, 42)]:
        m, x, dout, ref, ref_dx, ref_grads = _bilstm_case(cuda, T, N, I, 256, O, seed=seed)
        out, dx, grads = _bilstm_run(m, x, dout, "seq")
        assert int(crnn_engine.LAST_LSTM_FLAGS[-1]) == 0, "inter-CTA wait timed out"
        torch.testing.assert_close(out, ref, rtol=5e-2, atol=5e-2)
        # relative Frobenius error: bf16 operands, fp32 accumulation and state
        assert float((out - ref).norm() / ref.norm()) < 1e-2
        assert float((dx - ref_dx).norm() / ref_dx.norm()) < 2e-2
        for got, want in zip(grads, ref_grads):
            assert float((got - want).norm() / want.norm()) < 2e-2
