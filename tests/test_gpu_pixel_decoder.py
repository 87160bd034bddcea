"""GPU parity tests of the pixel decoder before the path (SURVEY.md section 8 row f3;
open_set/models/mask2former_head.py:787 -> mmdet MSDeformAttnPixelDecoder, configs/instance/coco_b48n17.py:38-70): every
stage kernel against the oracle's / torch's fp32 CPU arithmetic, forward and backward, then the whole module (fp32 FMA
mode: tight; tf32 tcgen05 mode: BASELINE.json's 1e-2 of range) and the module chained into the decoder head.  All calls go
through the C-ABI (ctypes -> libcgg_b200.so)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from cgg_b200 import synth
from cgg_b200.pixel_decoder import (build_pixel_decoder_from_state_dict, _PDRuntime, _Conv3x3, _GroupNorm, _MSDeformCore,
                                    _UpsampleAdd, _LevelToNCHW, _Conv1x1NCHW, _Conv1x1ToNCHW)
from cgg_b200.train import _K
from oracle import pixel_decoder_oracle as P

pytestmark = pytest.mark.gpu
DEV = 'cuda'
CHS = (32, 64, 96, 160)


@pytest.fixture(scope='module')
def rt():
    return _PDRuntime(torch.device(DEV, 0))


def _rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12)


@pytest.mark.parametrize('shapes', [[(4, 5), (8, 10), (16, 20)], [(3, 3), (6, 7), (12, 13)]])
def test_ms_deform_attn_forward_and_backward(rt, shapes):
    """mmcv MultiScaleDeformableAttention core vs its pure-torch twin (grid_sample), incl. taps outside every level."""
    k = _K(rt)
    g = torch.Generator().manual_seed(11)
    B, heads, L, Pn = 2, 8, 3, 4
    S = sum(h * w for h, w in shapes)
    value = torch.randn((B, S, 256), generator=g, requires_grad=True)
    off = (torch.randn((B, S, heads * L * Pn * 2), generator=g) * 3.0).requires_grad_(True)      # several pixels wide: leaves the maps
    lg = torch.randn((B, S, heads * L * Pn), generator=g, requires_grad=True)
    ref = P.reference_points(shapes)
    norm = torch.tensor([[w, h] for (h, w) in shapes], dtype=torch.float32)
    loc = ref[None, :, None, None, None, :] + off.view(B, S, heads, L, Pn, 2) / norm[None, None, None, :, None, :]
    aw = lg.view(B, S, heads, L * Pn).softmax(-1).view(B, S, heads, L, Pn)
    want = P.ms_deform_attn_core(value.view(B, S, heads, 32), shapes, loc, aw)
    probe = torch.randn(want.shape, generator=g)
    (want * probe).sum().backward()
    v, o, l_ = (t.detach().to(DEV).requires_grad_(True) for t in (value, off, lg))
    got = _MSDeformCore.apply(k, v, o, l_, shapes, heads, Pn)
    (got * probe.to(DEV)).sum().backward()
    assert _rel(got.detach().cpu(), want.detach()) < 2e-5
    assert _rel(v.grad.cpu(), value.grad) < 2e-5
    assert _rel(l_.grad.cpu(), lg.grad) < 5e-5
    assert _rel(o.grad.cpu(), off.grad) < 5e-5


@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('shape', [(2, 300, 256), (3, 1000, 64), (1, 7, 128)])
def test_group_norm_tokens_forward_and_backward(rt, relu, shape):
    k = _K(rt)
    B, Pn, Cc = shape
    G = 32 if Cc >= 128 else 16
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(shape, generator=g) * 2 + 0.7).requires_grad_(True)
    ga = (1 + 0.2 * torch.randn(Cc, generator=g)).requires_grad_(True)
    be = (0.1 * torch.randn(Cc, generator=g)).requires_grad_(True)
    want = F.group_norm(x.transpose(1, 2), G, ga, be, 1e-5).transpose(1, 2)
    if relu:
        want = torch.relu(want)
    probe = torch.randn(shape, generator=g)
    (want * probe).sum().backward()
    xs, gs, bs = (t.detach().to(DEV).requires_grad_(True) for t in (x, ga, be))
    got = _GroupNorm.apply(k, xs, gs, bs, G, relu)
    (got * probe.to(DEV)).sum().backward()
    assert _rel(got.detach().cpu(), want.detach()) < 1e-5
    assert _rel(xs.grad.cpu(), x.grad) < 5e-5
    assert _rel(gs.grad.cpu(), ga.grad) < 5e-5 and _rel(bs.grad.cpu(), be.grad) < 5e-5


@pytest.mark.parametrize('tf32', [False, True])
@pytest.mark.parametrize('geom', [(2, 12, 20, 64, 64), (1, 9, 140, 32, 96), (2, 5, 7, 256, 256)])
def test_conv3x3_implicit_gemm_forward_and_backward(rt, tf32, geom):
    """3x3 conv as an implicit GEMM (TMA zero-fill = the padding) against F.conv2d, with integer-valued operands so that the
    tf32 form must be exact too; includes a width that is not a multiple of the 128-row tile."""
    k = _K(rt, tf32=tf32)
    B, H, W, Cin, Cout = geom
    g = torch.Generator().manual_seed(3)
    x = torch.randint(-4, 5, (B, Cin, H, W), generator=g).float().requires_grad_(True)
    w = torch.randint(-3, 4, (Cout, Cin, 3, 3), generator=g).float().requires_grad_(True)
    want = F.conv2d(x, w, padding=1)
    probe = torch.randint(-2, 3, want.shape, generator=g).float()
    (want * probe).sum().backward()
    xt = x.detach().permute(0, 2, 3, 1).reshape(B, H * W, Cin).contiguous().to(DEV).requires_grad_(True)
    w2 = w.detach().permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous().to(DEV).requires_grad_(True)
    got = _Conv3x3.apply(k, xt, w2, H, W)
    (got * probe.permute(0, 2, 3, 1).reshape(B, H * W, Cout).to(DEV)).sum().backward()
    assert torch.equal(got.detach().cpu().view(B, H, W, Cout).permute(0, 3, 1, 2), want.detach())
    assert torch.equal(xt.grad.cpu().view(B, H, W, Cin).permute(0, 3, 1, 2), x.grad)
    assert torch.equal(w2.grad.cpu().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2), w.grad)


def test_upsample_add_and_layout_changes(rt):
    k = _K(rt)
    g = torch.Generator().manual_seed(9)
    B, Cc, (h, w), (H, W) = 2, 64, (6, 9), (12, 18)
    S = 10 + h * w
    tok = torch.randn((B, S, Cc), generator=g, requires_grad=True)
    lat = torch.randn((B, H * W, Cc), generator=g, requires_grad=True)
    coarse = tok[:, 10:].transpose(1, 2).reshape(B, Cc, h, w)
    want = lat + F.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=False).flatten(2).transpose(1, 2)
    probe = torch.randn(want.shape, generator=g)
    (want * probe).sum().backward()
    ts, ls = tok.detach().to(DEV).requires_grad_(True), lat.detach().to(DEV).requires_grad_(True)
    got = _UpsampleAdd.apply(k, ls, ts, 10, (h, w), (H, W))
    (got * probe.to(DEV)).sum().backward()
    assert _rel(got.detach().cpu(), want.detach()) < 1e-6
    assert _rel(ts.grad.cpu(), tok.grad) < 1e-5 and torch.equal(ls.grad.cpu(), lat.grad)
    # tokens -> NCHW (fp32 and bf16) and its adjoint
    t2 = tok.detach().to(DEV).requires_grad_(True)
    out = _LevelToNCHW.apply(k, t2, 10, (h, w), False)
    assert torch.equal(out.detach().cpu(), coarse.detach())
    pr = torch.randn(out.shape, generator=g)
    (out * pr.to(DEV)).sum().backward()
    wantg = torch.zeros(B, S, Cc)
    wantg[:, 10:] = pr.flatten(2).transpose(1, 2)
    assert torch.equal(t2.grad.cpu(), wantg)
    with torch.no_grad():
        ob = _LevelToNCHW.apply(k, t2.detach(), 10, (h, w), True)
    assert ob.dtype == torch.bfloat16 and torch.equal(ob.cpu(), coarse.detach().bfloat16())


@pytest.mark.parametrize('tf32', [False, True])
def test_conv1x1_nodes(rt, tf32):
    k = _K(rt, tf32=tf32)
    g = torch.Generator().manual_seed(21)
    B, Cin, N, h, w = 2, 96, 64, 8, 12
    x = torch.randint(-4, 5, (B, Cin, h, w), generator=g).float().requires_grad_(True)
    W = torch.randint(-3, 4, (N, Cin), generator=g).float().requires_grad_(True)
    b = torch.randint(-3, 4, (N,), generator=g).float().requires_grad_(True)
    want = F.conv2d(x, W.view(N, Cin, 1, 1), b)
    probe = torch.randint(-2, 3, want.shape, generator=g).float()
    (want * probe).sum().backward()
    xs, Ws, bs = (t.detach().to(DEV).requires_grad_(True) for t in (x, W, b))
    got = _Conv1x1NCHW.apply(k, xs, Ws, bs)                                   # NCHW in, tokens out
    (got * probe.flatten(2).transpose(1, 2).to(DEV)).sum().backward()
    assert torch.equal(got.detach().cpu().transpose(1, 2).reshape(B, N, h, w), want.detach())
    assert torch.equal(xs.grad.cpu(), x.grad) and torch.equal(Ws.grad.cpu(), W.grad) and torch.equal(bs.grad.cpu(), b.grad)
    xt = x.detach().flatten(2).transpose(1, 2).contiguous().to(DEV).requires_grad_(True)
    Ws2, bs2 = Ws.detach().clone().requires_grad_(True), bs.detach().clone().requires_grad_(True)
    got2 = _Conv1x1ToNCHW.apply(k, xt, Ws2, bs2, h, w)                        # tokens in, NCHW out
    (got2 * probe.to(DEV)).sum().backward()
    assert torch.equal(got2.detach().cpu(), want.detach())
    assert torch.equal(xt.grad.cpu().transpose(1, 2).reshape(B, Cin, h, w), x.grad)
    assert torch.equal(Ws2.grad.cpu(), W.grad) and torch.equal(bs2.grad.cpu(), b.grad)


def _oracle(sd, feats, grad=False):
    if not grad:
        with torch.no_grad():
            return P.pixel_decoder_forward(sd, feats)
    return P.pixel_decoder_forward(sd, feats)


@pytest.mark.parametrize('precision,tol', [('fp32', 2e-4), ('tf32', 1e-2)])
@pytest.mark.parametrize('size', [(2, 128, 160), (1, 96, 224)])
def test_pixel_decoder_forward_matches_the_oracle(precision, tol, size):
    B, H, W = size
    sd = synth.make_pixel_decoder_params(2, in_channels=CHS)
    feats = synth.make_backbone_feats(2, B, H, W, CHS)
    mf, mems = _oracle(sd, feats)
    m = build_pixel_decoder_from_state_dict(sd, CHS, DEV, precision=precision).eval()
    with torch.no_grad():
        got_mf, got_mems = m([f.to(DEV) for f in feats])
    torch.cuda.synchronize()
    assert got_mf.shape == mf.shape and [t.shape for t in got_mems] == [t.shape for t in mems]
    assert _rel(got_mf.cpu(), mf) < tol, _rel(got_mf.cpu(), mf)
    for a, b in zip(got_mems, mems):
        assert _rel(a.cpu(), b) < tol, _rel(a.cpu(), b)


def test_pixel_decoder_gradients_match_the_oracle_autograd():
    """fp32 mode: the gradient of every state_dict key and of every backbone map against the oracle's autograd."""
    B, H, W = 2, 96, 128
    sd = synth.make_pixel_decoder_params(4, in_channels=CHS)
    feats = synth.make_backbone_feats(4, B, H, W, CHS)
    sd_o = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    feats_o = [f.clone().requires_grad_(True) for f in feats]
    mf, mems = _oracle(sd_o, feats_o, grad=True)
    g = torch.Generator().manual_seed(8)
    probes = [torch.randn(t.shape, generator=g) for t in [mf] + list(mems)]
    sum((t * p).sum() for t, p in zip([mf] + list(mems), probes)).backward()
    m = build_pixel_decoder_from_state_dict(sd, CHS, DEV, precision='fp32').train()
    fs = [f.to(DEV).requires_grad_(True) for f in feats]
    got_mf, got_mems = m(fs)
    sum((t * p.to(DEV)).sum() for t, p in zip([got_mf] + list(got_mems), probes)).backward()
    torch.cuda.synchronize()
    bad = []
    for name, p in m.named_parameters():
        want = sd_o[name].grad
        assert p.grad is not None, name
        e = _rel(p.grad.cpu(), want)
        if e > 2e-3:
            bad.append((name, e))
    for i, (f, fo) in enumerate(zip(fs, feats_o)):
        e = _rel(f.grad.cpu(), fo.grad)
        if e > 2e-3:
            bad.append(('feats[%d]' % i, e))
    assert not bad, bad


def test_pixel_decoder_feeds_the_decoder_head():
    """Mask2FormerHeadOpenB200.forward(feats, img_metas) with the B200 pixel decoder attached (head.py:787 onward) equals
    oracle(pixel decoder) -> oracle(decoder head), fp32 parity mode."""
    from cgg_b200.head import build_head_from_state_dict
    from oracle import cgg_oracle as O
    chs = (32, 64, 96, 160)
    B, H, W, Q = 1, 128, 160, 20
    sd_p = synth.make_pixel_decoder_params(6, in_channels=chs)
    sd_h = synth.make_params(seed=6, num_queries=Q, perturb=True)
    feats = synth.make_backbone_feats(6, B, H, W, chs)
    with torch.no_grad():
        mf, mems = P.pixel_decoder_forward(sd_p, feats)
        ref = O.decoder_forward(sd_h, mf, mems)
    head = build_head_from_state_dict(sd_h, Q, 49, 'fp32', DEV)
    head.pixel_decoder = build_pixel_decoder_from_state_dict(sd_p, chs, DEV, precision='fp32').eval()
    with torch.no_grad():
        cls, emb, mask = head([f.to(DEV) for f in feats], [dict()] * B)
    torch.cuda.synchronize()
    for j in (0, 9):
        assert _rel(mask[j].float().cpu(), ref['mask'][j]) < 2e-3
        assert _rel(cls[j].float().cpu(), ref['cls'][j]) < 2e-3


def test_training_chain_backbone_maps_to_head_losses():
    """The whole trained module of the reference's head: backbone maps -> pixel decoder -> decoder head -> a loss over all
    ten head calls -> gradients of the pixel decoder's parameters, the head's parameters and the backbone maps, against the
    oracle chain's autograd (fp32 FMA mode; attention masks forced to the oracle's, they are detached constants of the graph)."""
    from cgg_b200.head import build_head_from_state_dict
    from cgg_b200.train import decoder_forward_train
    from oracle import cgg_oracle as O
    chs = (32, 64, 96, 160)
    B, H, W, Q = 1, 96, 128, 12
    sd_p = synth.make_pixel_decoder_params(9, in_channels=chs)
    sd_h = synth.make_params(seed=9, num_queries=Q, perturb=True)
    feats = synth.make_backbone_feats(9, B, H, W, chs)
    g = torch.Generator().manual_seed(9)
    sd_po = {k: v.clone().requires_grad_(True) for k, v in sd_p.items()}
    sd_ho = {k: v.clone().requires_grad_(k != 'class_embs') for k, v in sd_h.items()}
    feats_o = [f.clone().requires_grad_(True) for f in feats]
    mf_o, mems_o = P.pixel_decoder_forward(sd_po, feats_o)
    ref = O.decoder_forward(sd_ho, mf_o, mems_o)
    probes = [torch.randn(ref['mask'][j].shape, generator=g) for j in range(10)]
    eprobes = [torch.randn(ref['emb'][j].shape, generator=g) for j in range(10)]
    sum((ref['mask'][j] * probes[j]).sum() * 0.01 + (ref['emb'][j] * eprobes[j]).sum() * 0.1 for j in range(10)).backward()
    pd = build_pixel_decoder_from_state_dict(sd_p, chs, DEV, precision='fp32').train()
    head = build_head_from_state_dict(sd_h, Q, 49, 'fp32', DEV, train_precision='fp32').train()
    fs = [f.to(DEV).requires_grad_(True) for f in feats]
    mf, mems = pd(fs)
    assert mf.requires_grad and mems[0].requires_grad
    forced = [(O.pack_mask_bits(ref['masked'][j].detach()).to(DEV), ref['masked'][j].detach().all(-1).to(torch.uint8).to(DEV))
              for j in range(9)]
    cls, emb, mask = decoder_forward_train(head, mf, mems, forced_attn_masks=forced)
    sum((mask[j] * probes[j].to(DEV)).sum() * 0.01 + (emb[j] * eprobes[j].to(DEV)).sum() * 0.1 for j in range(10)).backward()
    torch.cuda.synchronize()
    assert _rel(mask[9].detach().cpu(), ref['mask'][9].detach()) < 2e-3
    bad = [(n, _rel(p.grad.cpu(), sd_po[n].grad)) for n, p in pd.named_parameters() if _rel(p.grad.cpu(), sd_po[n].grad) > 5e-3]
    bad += [('feats[%d]' % i, _rel(f.grad.cpu(), fo.grad)) for i, (f, fo) in enumerate(zip(fs, feats_o))
            if _rel(f.grad.cpu(), fo.grad) > 5e-3]
    for n, p in head.named_parameters():
        if sd_ho[n].grad is None or float(sd_ho[n].grad.abs().max()) == 0.0:      # (the class logits are not in this loss)
            continue
        if _rel(p.grad.cpu(), sd_ho[n].grad) > 5e-3:
            bad.append((n, _rel(p.grad.cpu(), sd_ho[n].grad)))
    assert not bad, bad


@pytest.mark.parametrize('precision,tol', [('fp32', 2e-4), ('tf32', 1e-2)])
def test_pixel_decoder_matches_the_committed_golden_vectors(precision, tol):
    """The CUDA path against tests/golden/pixel_decoder.npz: outputs of an independent implementation (HF's pixel decoder)."""
    import os
    import numpy as np
    import make_pixel_decoder_golden as G
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(G.__file__)), 'pixel_decoder.npz'))
    c = G.CASE
    sd = synth.make_pixel_decoder_params(c['seed'], in_channels=c['in_channels'])
    feats = synth.make_backbone_feats(c['seed'], c['batch'], c['height'], c['width'], c['in_channels'])
    m = build_pixel_decoder_from_state_dict(sd, c['in_channels'], DEV, precision=precision).eval()
    with torch.no_grad():
        mf, mems = m([f.to(DEV) for f in feats])
    torch.cuda.synchronize()
    assert _rel(mf.cpu(), torch.from_numpy(z['mask_features'])) < tol
    for i, t in enumerate(mems):
        assert _rel(t.cpu(), torch.from_numpy(z['memory%d' % i])) < tol
