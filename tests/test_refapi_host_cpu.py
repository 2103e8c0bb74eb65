"""CPU: host-side behaviour of the reference-named surfaces that needs no GPU: constructor signatures / state-dict keys,
the clamp constant of CTCDecoder2D following its `saved_tiny` buffer, and the loud refusal of CPU tensors."""
import pytest
import torch


_TOP = ("ops", "decoders", "backbones", "assets", "config", "concern")


@pytest.fixture(scope="module")
def api():
    """The reference-named top-level packages resolved to megreader_b200/refapi for this module only: whatever another test
    module imported under the same names (the reference itself, through oracle/ref_loader.py) is put back afterwards."""
    import os
    import sys
    import megreader_b200
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _TOP}
    for k in saved:
        del sys.modules[k]
    old_path = list(sys.path)
    refapi_dir = os.path.join(os.path.dirname(megreader_b200.__file__), "refapi")
    sys.path.insert(0, refapi_dir)
    try:
        import backbones
        import decoders
        import assets.ops.dcn  # noqa: F401
        assert "refapi" in decoders.__file__ and "refapi" in backbones.__file__
        yield backbones, decoders
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _TOP]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = old_path


def test_ctc_decoder2d_tiny_follows_the_buffer(api):
    _, decoders = api
    dec = decoders.CTCDecoder2D(16, inner_channels=8)
    assert sorted(dec.state_dict()) == ["pred_classify.1.bias", "pred_classify.1.weight", "pred_classify.2.bias",
                                        "pred_classify.2.weight", "pred_mask.1.bias", "pred_mask.1.weight",
                                        "pred_mask.2.bias", "pred_mask.2.weight", "saved_tiny"]
    assert dec._tiny() == float(torch.finfo(torch.float32).tiny)
    dec.saved_tiny.fill_(0.25)                                   # in-place change (what load_state_dict does)
    assert dec._tiny() == 0.25
    sd = dec.state_dict()
    sd["saved_tiny"] = torch.tensor(0.5)
    dec.load_state_dict(sd)
    assert dec._tiny() == 0.5
    dec = dec.double().float()                                   # buffer object replaced by _apply
    assert dec._tiny() == 0.5


def test_training_surfaces_refuse_cpu(api):
    backbones, decoders = api
    dec2d = decoders.CTCDecoder2D(16, inner_channels=8).train()
    with pytest.raises(NotImplementedError):
        dec2d(torch.zeros(1, 16, 4, 8), targets=torch.zeros(1, 32), lengths=torch.ones(1), train=True)
    with pytest.raises(NotImplementedError):
        backbones.crnn_backbone()(torch.zeros(1, 3, 32, 32))
    ctc = decoders.CTCDecoder(16, inner_channels=8).train()
    with pytest.raises(NotImplementedError):
        ctc(torch.zeros(1, 16, 16, 64), targets=torch.zeros(1, 32, dtype=torch.long), lengths=torch.ones(1, dtype=torch.long),
            train=True)


def test_public_names_match_the_reference_inits(api):
    backbones, decoders = api
    for name in ("Resnet18FPN", "Resnet34FPN", "Resnet50FPN", "Resnet101FPN", "Resnet152FPN", "resnet18", "resnet34",
                 "resnet50", "resnet101", "deformable_resnet50", "crnn_backbone", "resnet50dilated_ppm"):
        assert callable(getattr(backbones, name)), name           # backbones/__init__.py:1-4
    for name in ("AttentionDecoder", "CTCDecoder2D", "CTCDecoder", "CRNNDecoder", "CTCLoss2D", "CTC2DLoss", "EASTDecoder"):
        assert callable(getattr(decoders, name)), name
    import assets.ops.dcn as dcn
    assert sorted(dcn.__all__) == sorted(['DeformConv', 'DeformConvPack', 'ModulatedDeformConv', 'ModulatedDeformConvPack',
                                          'DeformRoIPooling', 'DeformRoIPoolingPack', 'ModulatedDeformRoIPoolingPack',
                                          'deform_conv', 'modulated_deform_conv', 'deform_roi_pooling'])


def test_engine_modules_refuse_cpu_tensors():
    """the modules use_engine_convs() installs have no CPU fallback: they raise on CPU tensors, and restore_library_convs() gives the
    framework modules back (same parameter objects)"""
    import pytest
    import torch
    from megreader_b200 import conv_engine
    m = torch.nn.Sequential(torch.nn.Conv2d(64, 64, 1), torch.nn.BatchNorm2d(64), torch.nn.ConvTranspose2d(64, 32, 2, 2))
    params = [id(p) for p in m.parameters()]
    assert conv_engine.use_engine_convs(m) == 1
    assert [type(x) for x in m] == [conv_engine.EngineConv2d, conv_engine.EngineBatchNorm2d, conv_engine.EngineConvTranspose2d]
    x = torch.zeros(1, 64, 4, 4)
    for layer in m:
        with pytest.raises(NotImplementedError):
            layer(x)
    assert conv_engine.restore_library_convs(m) == 1
    assert [type(x) for x in m] == [torch.nn.Conv2d, torch.nn.BatchNorm2d, torch.nn.ConvTranspose2d]
    assert [id(p) for p in m.parameters()] == params
    assert tuple(m(x).shape) == (1, 32, 8, 8)


def test_attention_feedback_draws_follow_the_reference_order():
    """AttentionDecoder.draw_feedback makes the training loop's random draws before the loop: per step the numpy teacher-forcing coin
    (attention_decoder.py:51-54, :107), then torch.rand and torch.randint of the step dropout (:112-116) -- the same generators in the
    same order as the reference's in-loop draws, so seeded runs line up"""
    import numpy as np
    import torch
    import megreader_b200.refapi.decoders as md
    m = md.AttentionDecoder(32, inner_channels=64, max_size=8, height=1, step_dropout=0.3)
    n, vocab = 5, len(m.charset)
    np.random.seed(11)
    torch.manual_seed(11)
    coin, swap, noise = m.draw_feedback(n)
    np.random.seed(11)
    torch.manual_seed(11)
    for t in range(m.max_size):                      # the reference's loop body, draws only
        c = np.random.rand() < 0.5
        f = (torch.rand(n) < 0.3).long()
        r = torch.randint(high=vocab, size=(n,))
        assert bool(coin[t]) == bool(c) and torch.equal(swap[t], f) and torch.equal(noise[t], r)
    assert coin.dtype == torch.bool and tuple(swap.shape) == (8, n) == tuple(noise.shape)
    m2 = md.AttentionDecoder(32, inner_channels=64, max_size=8, height=1, gt_as_output=True)
    assert bool(m2.draw_feedback(3)[0].all()) and int(m2.draw_feedback(3)[1].sum()) == 0
