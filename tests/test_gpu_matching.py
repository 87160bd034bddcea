"""GPU tests of the matching-based losses (cgg_b200/matching.py + csrc/match_kernels.cu) against the
oracle (oracle/matching_oracle.py, pinned on the live reference's loss_single) and the reference-generated fixture
tests/golden/matching.npz.  The random point draws are fed from the same seeded CPU torch.rand stream on both sides."""
import os
import numpy as np
import pytest
import torch

from oracle import matching_oracle as MO
from cgg_b200 import synth, matching
from cgg_b200.head import build_head_from_state_dict
from test_matching_cpu import make_case

pytestmark = pytest.mark.gpu
DEV = 'cuda'
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def head():
    return build_head_from_state_dict(synth.make_params(seed=2, num_queries=20), 20, 49, 'fp32', DEV)


def _ml(head, P):
    ml = matching.MatchingLosses(head, train_cfg=dict(num_points=P))
    ml.rand = lambda *shape, device=None: torch.rand(*shape)          # the oracle's stream: seeded CPU generator
    return ml


def test_point_sample_forward_backward_match_grid_sample():
    g = torch.Generator().manual_seed(0)
    inp = torch.randn((6, 37, 53), generator=g)
    coords = torch.rand((6, 900, 2), generator=g) * 1.1 - 0.05          # a few points outside [0, 1]: zero padding
    want_in = inp.clone().requires_grad_(True)
    want = MO.point_sample(want_in.unsqueeze(1), coords).squeeze(1)
    dout = torch.randn(want.shape, generator=g)
    want.backward(dout)
    x = inp.to(DEV).requires_grad_(True)
    got = matching.point_sample(x, coords.to(DEV))
    got.backward(dout.to(DEV))
    assert float((got.detach().cpu() - want.detach()).abs().max()) < 1e-5
    assert float((x.grad.cpu() - want_in.grad).abs().max()) < 1e-4
    # shared coordinates (the matching step): one point set for every plane
    got1 = matching.point_sample(inp.to(DEV), coords[:1].to(DEV))
    want1 = MO.point_sample(inp.unsqueeze(1), coords[:1].repeat(6, 1, 1)).squeeze(1)
    assert float((got1.cpu() - want1).abs().max()) < 1e-5


@pytest.mark.parametrize('G', [1, 5, 17])
def test_matching_cost_matches_oracle(G):
    g = torch.Generator().manual_seed(G)
    Q, P, C1 = 20, 777, 49
    x = torch.randn((Q, P), generator=g) * 3
    gt = torch.rand((G, P), generator=g).round() * torch.rand((G, P), generator=g).clamp(min=0.3)   # fractional samples too
    cls, emb = torch.randn((Q, C1), generator=g), torch.randn((Q, C1), generator=g) * 2
    labels = torch.randint(0, C1 - 1, (G,), generator=g)
    w = dict(cls=0.7, cls_emb=2.0, mask=5.0, dice=5.0, dice_eps=1.0)
    want = MO.matching_cost(cls, emb, x, labels, gt, w)
    got = matching.matching_cost(x.to(DEV), gt.to(DEV), labels.to(DEV), cls.to(DEV), emb.to(DEV), 0.7, 2.0, 5.0, 5.0, 1.0)
    assert float((got.cpu() - want).abs().max()) < 2e-5 * float(want.abs().max())


@pytest.mark.parametrize('seed,gts', [(1, (3, 5)), (2, (1, 0)), (3, (7, 2)), (4, (0, 0))])
def test_loss_single_matches_oracle_with_gradients(head, seed, gts):
    ncls, P = 48, 512
    cls_scores, cls_emb_logits, mask_preds, gt_labels, gt_masks = make_case(seed, gts=gts)
    a = [t.clone().requires_grad_(True) for t in (cls_scores, cls_emb_logits, mask_preds)]
    torch.manual_seed(50 + seed)
    want = MO.loss_single_matching(a[0], a[1], a[2], gt_labels, gt_masks, ncls, dict(num_points=P, loss_cls_weight=0.5))
    total_w = want['loss_cls'] + want['loss_cls_emb'] + want['loss_mask'] + want['loss_dice']
    total_w.backward()
    ml = _ml(head, P)
    ml.loss_cls_weight = 0.5
    b = [t.clone().to(DEV).requires_grad_(True) for t in (cls_scores, cls_emb_logits, mask_preds)]
    torch.manual_seed(50 + seed)
    got, labels, mask_weights = ml.loss_single(b[0], b[1], b[2], gt_labels, gt_masks)
    assert torch.equal(labels.cpu(), want['labels']) and torch.equal(mask_weights.cpu(), want['mask_weights'])
    for k in ('loss_cls', 'loss_cls_emb', 'loss_mask', 'loss_dice'):
        assert abs(float(got[k]) - float(want[k])) < 2e-5 * max(1.0, abs(float(want[k]))), (k, float(got[k]), float(want[k]))
    (got['loss_cls'] + got['loss_cls_emb'] + got['loss_mask'] + got['loss_dice']).backward()
    for x, y, name in zip(b, a, ('cls_scores', 'cls_emb_logits', 'mask_preds')):
        assert x.grad is not None and y.grad is not None, name
        err = float((x.grad.cpu() - y.grad).abs().max())
        assert err < 1e-4 * float(y.grad.abs().max()) + 1e-9, (name, err)


def test_reference_fixture(head):
    z = np.load(os.path.join(HERE, 'golden', 'matching.npz'))
    B = int(z['B'])
    gt_labels = [torch.from_numpy(z['gt_labels_%d' % b]) for b in range(B)]
    gt_masks = [torch.from_numpy(z['gt_masks_%d' % b]) for b in range(B)]
    ml = _ml(head, int(z['num_points']))
    torch.manual_seed(int(z['seed']))
    got, labels, _ = ml.loss_single(torch.from_numpy(z['cls_scores']).to(DEV), torch.from_numpy(z['cls_emb_logits']).to(DEV),
                                    torch.from_numpy(z['mask_preds']).to(DEV), gt_labels, gt_masks)
    assert torch.equal(labels.cpu(), torch.from_numpy(z['labels']))
    for k in ('loss_cls', 'loss_cls_emb', 'loss_mask', 'loss_dice'):
        assert abs(float(got[k]) - float(z[k])) < 2e-5 * max(1.0, abs(float(z[k]))), (k, float(got[k]), float(z[k]))


def test_forward_train_returns_the_reference_loss_dict():
    """forward_train (head.py:851-921) with the matching losses attached: the reference's loss keys for all 10 head calls
    (caption generation excepted), every parameter that the losses depend on receives a gradient."""
    from test_gpu_post import _caption_head
    Q, B, ncls1 = 16, 2, 49
    head, sd, _, _ = _caption_head(Q)
    head.train()
    head.matching_losses = matching.MatchingLosses(head, train_cfg=dict(num_points=256))
    mf, mems = synth.make_inputs(9, B, 96, 128)
    feats = [mf.to(DEV)] + [m.to(DEV) for m in mems]
    g = torch.Generator().manual_seed(1)
    H4, W4 = mf.shape[-2:]
    gt_labels = [torch.randint(0, ncls1 - 1, (3,), generator=g).to(DEV), torch.randint(0, ncls1 - 1, (2,), generator=g).to(DEV)]
    gt_masks = [(torch.rand((3, H4, W4), generator=g) > 0.7).to(DEV), (torch.rand((2, H4, W4), generator=g) > 0.7).to(DEV)]
    ids = torch.randint(1000, 1500, (B, 35), generator=g).to(DEV)
    cmask = torch.zeros((B, 35), dtype=torch.long, device=DEV)
    cmask[0, :4] = 1
    cmask[1, :9] = 1
    losses = head.forward_train(feats, [dict() for _ in range(B)], None, gt_labels, gt_masks, None, list(ids), list(cmask),
                                list(ids), list(cmask))
    want = {p + k for p in [''] + ['d%d.' % i for i in range(9)] for k in ('loss_cls', 'loss_cls_emb', 'loss_mask', 'loss_dice', 'loss_grounding')}
    assert set(losses) == want
    sum(losses.values()).backward()
    missing = [k for k, p in head.named_parameters() if p.requires_grad and p.grad is None]
    assert missing == ['cls_embed.weight', 'cls_embed.bias'] or missing == [], missing     # loss_cls has weight 0 in this config
    assert all(torch.isfinite(v) for v in losses.values())


def test_simple_test_assigns_labels_like_get_target_single():
    """simple_test(gt_labels=, gt_masks=) (head.py:947-953): assigned_labels = the Hungarian assignment of image 0."""
    from test_gpu_post import _caption_head
    from oracle import cgg_oracle as O
    Q, B = 24, 1
    head, sd, _, _ = _caption_head(Q)
    head.eval()
    head.matching_losses = _ml(head, 300)
    mf, mems = synth.make_inputs(12, B, 128, 160)
    feats = [mf.to(DEV)] + [m.to(DEV) for m in mems]
    g = torch.Generator().manual_seed(3)
    gt_labels = torch.randint(0, 48, (4,), generator=g)
    gt_masks = (torch.rand((4, 128, 160), generator=g) > 0.6).long()            # full resolution, the logits are 1/4 scale
    metas = [dict(batch_input_shape=(128, 160), pad_shape=(128, 160, 3))]
    torch.manual_seed(77)
    with torch.no_grad():
        got = head.simple_test(feats, metas, gt_labels=[[gt_labels]], gt_masks=[[gt_masks]])[0]
    ref = O.decoder_forward({k: v for k, v in sd.items() if not k.startswith('bert')}, mf, mems)
    logits = O.cls_emb_logits(ref['emb'][9][0], sd['class_embs'], 10.0)
    torch.manual_seed(77)
    want = MO.get_target_single(ref['cls'][9][0], logits, ref['mask'][9][0], gt_labels, gt_masks, 48,
                                dict(MO.DEFAULT_CFG, num_points=300))[0]
    assert torch.equal(got.cpu(), want)
