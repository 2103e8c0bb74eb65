"""Shared checker: megreader_b200.refapi surfaces (resnet trunks, PPM, FPN, attention head, 1-D CTC conv head) against
tests/golden/surfaces_ref.npz, which oracle/make_golden.py produced from the UNMODIFIED reference modules on CPU."""
import os

import numpy as np
import torch

from tests.weights import fill_state_dict, surface_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden", "surfaces_ref.npz")


def close(a, ref, what, tol):
    a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    scale = max(1e-6, float(np.abs(ref).max()))
    np.testing.assert_allclose(a, ref, rtol=tol, atol=tol * scale, err_msg=what)


def gnorms(module, g, prefix, tol):
    params = dict(module.named_parameters())
    for key in g.files:
        if key.startswith(prefix + ".gnorm."):
            name = key[len(prefix) + 7:]
            got = params[name].grad.double().norm().item()
            np.testing.assert_allclose(got, float(g[key]), rtol=max(tol, 1e-4) * 10, err_msg=key)


def check_backbones(device, tol):
    import megreader_b200.refapi.backbones as mb
    g = np.load(GOLD)
    x, x2, _, _, _ = surface_inputs()
    tx = torch.from_numpy(x).to(device)
    with torch.no_grad():
        m = fill_state_dict(mb.resnet18(pretrained=False), "r18.").to(device).eval()
        for i, f in enumerate(m(tx)):
            close(f, g["r18.%d" % i], "resnet18 stage %d" % i, tol)
        m = fill_state_dict(mb.resnet50dilated_ppm(), "ppm.").to(device).eval()
        close(m(tx), g["ppm"], "resnet50dilated_ppm", tol)
        m = fill_state_dict(mb.Resnet50FPN(resnet_pretrained=False), "fpn50.").to(device).eval()
        close(m(tx), g["fpn50"], "Resnet50FPN", tol)
    m = fill_state_dict(mb.Resnet18FPN(resnet_pretrained=False), "fpn18.").to(device).train()
    y = m(torch.from_numpy(x2).to(device))
    y.square().mean().backward()
    close(y, g["fpn18.train"], "Resnet18FPN train", tol)
    gnorms(m, g, "fpn18", tol)


def check_attention(device, tol):
    import megreader_b200.refapi.decoders as md
    g = np.load(GOLD)
    _, _, feat, targets, lengths = surface_inputs()
    tf, tt, tl = (torch.from_numpy(a).to(device) for a in (feat, targets, lengths))
    att = fill_state_dict(md.AttentionDecoder(256, gt_as_output=True), "attn.").to(device).train()
    loss, amap = att(tf, targets=tt, lengths=tl)
    loss.sum().backward()
    close(loss, g["attn.loss"], "attention loss", tol)
    close(amap, g["attn.map"], "attention maps", tol)
    gnorms(att, g, "attn", tol)
    with torch.no_grad():
        pred = att.eval()(tf)
    assert pred.dtype == torch.int32 and tuple(pred.shape) == (3, 32)
    assert np.array_equal(pred.cpu().numpy(), g["attn.eval"])


def check_ctc_head(device, tol, train):
    import megreader_b200.refapi.decoders as md
    g = np.load(GOLD)
    _, _, feat, targets, lengths = surface_inputs()
    tf, tt, tl = (torch.from_numpy(a).to(device) for a in (feat, targets, lengths))
    ctc = fill_state_dict(md.CTCDecoder(256), "ctc1d.").to(device)
    with torch.no_grad():                                     # eval first: pristine BN running statistics
        close(ctc.eval()(tf, train=False), g["ctc1d.eval"], "ctc head eval", tol)
    if train:
        loss, lp = ctc.train()(tf, targets=tt, lengths=tl, train=True)
        loss.backward()
        np.testing.assert_allclose(loss.item(), float(g["ctc1d.loss"]), rtol=max(tol, 1e-4))
        close(lp, g["ctc1d.log_probs"], "ctc head log-probs", tol)
        gnorms(ctc, g, "ctc1d", tol)
