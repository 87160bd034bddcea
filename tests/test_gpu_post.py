"""GPU tests of the step after the path at test time: the fused final-upsample /
instance-scoring kernels against the reference's own sequence of torch ops (mask2former_head.py:957-964,
maskformer_fusion_head.py:297-366, :412-425; mmdet mask2bbox restated below)."""
import pytest
import torch
import torch.nn.functional as F

from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
from cgg_b200 import postprocess as P

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _mask2bbox(masks):
    """mmdet.core.mask2bbox (mmdet 2.28)."""
    N = masks.shape[0]
    bboxes = masks.new_zeros((N, 4), dtype=torch.float32)
    x_any, y_any = torch.any(masks, dim=1), torch.any(masks, dim=2)
    for i in range(N):
        x, y = torch.where(x_any[i, :])[0], torch.where(y_any[i, :])[0]
        if len(x) > 0 and len(y) > 0:
            bboxes[i, :] = bboxes.new_tensor([x[0], y[0], x[-1] + 1, y[-1] + 1])
    return bboxes


def _reference_instance(mask_cls_emb, mask_pred_lowres, class_embs, meta, rescale, max_per_image=100):
    """simple_test upsample + fusion-head simple_test crop / rescale + instance_postprocess_emb, plain torch ops."""
    up = meta['batch_input_shape']
    mp = F.interpolate(mask_pred_lowres[None], size=up, mode='bilinear', align_corners=False)[0]
    mp = mp[:, :meta['img_shape'][0], :meta['img_shape'][1]]
    if rescale:
        mp = F.interpolate(mp[:, None], size=meta['ori_shape'][:2], mode='bilinear', align_corners=False)[:, 0]
    scores = F.softmax(mask_cls_emb @ class_embs.t(), -1)[:, :-1]
    nq, ncls = scores.shape
    sc, top = scores.flatten().topk(max_per_image, sorted=False)
    labels, qi = top % ncls, top // ncls
    m = mp[qi]
    binary = (m > 0).float()
    mscore = (m.sigmoid() * binary).flatten(1).sum(1) / (binary.flatten(1).sum(1) + 1e-6)
    return labels, torch.cat([_mask2bbox(binary.bool()), (sc * mscore)[:, None]], -1), binary.bool(), top


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('rescale', [False, True])
def test_fused_instance_postprocess_matches_reference_ops(dtype, rescale):
    Q, B = 100, 2
    sd = synth.make_params(seed=2, num_queries=Q)
    head = build_head_from_state_dict(sd, Q, 49, 'fp32', DEV)
    g = torch.Generator().manual_seed(9)
    h4, w4 = 64, 80                                   # padded input 256 x 320
    logits = (torch.randn((B, Q, h4, w4), generator=g) * 3).to(DEV).to(dtype)
    logits[0, 3] = -4.0                                # an empty mask (bbox zeros, score 0)
    logits[1, 5] = 4.0                                 # a full mask
    emb = torch.randn((B, Q, 768), generator=g).to(DEV) * 0.2
    ce = sd['class_embs'].to(DEV)
    metas = [dict(batch_input_shape=(256, 320), img_shape=(250, 300, 3), ori_shape=(375, 450, 3)),
             dict(batch_input_shape=(256, 320), img_shape=(256, 277, 3), ori_shape=(511, 553, 3))]
    got = P.instance_postprocess_emb_fused(head, emb, logits, ce, metas, rescale=rescale)
    for b in range(B):
        labels, boxes, masks, top = _reference_instance(emb[b], logits[b].float(), ce, metas[b], rescale)
        g_labels, g_boxes, g_bits, (H, W) = got[b]
        # top-k is unsorted in both: align by (query, class) id
        g_top = None
        order_ref = torch.argsort(top)
        sc_all = P.cls_emb_scores(head, emb[b:b + 1], ce)[0, :, :-1].flatten()
        g_top = sc_all.topk(100, sorted=False)[1]
        order_got = torch.argsort(g_top)
        assert torch.equal(top[order_ref], g_top[order_got])
        assert torch.equal(labels[order_ref], g_labels[order_got])
        g_masks = P.unpack_masks(g_bits, W)[order_got]
        assert g_masks.shape == masks.shape
        mism = (g_masks != masks[order_ref]).float().mean().item()
        assert mism < 1e-6, mism                       # `> 0` of two fp32 bilinear evaluations: bit-identical up to a stray ulp at 0
        torch.testing.assert_close(g_boxes[order_got][:, :4], boxes[order_ref][:, :4], rtol=0, atol=1.0 if mism > 0 else 0.0)
        torch.testing.assert_close(g_boxes[order_got][:, 4], boxes[order_ref][:, 4], rtol=2e-4, atol=1e-6)


def test_upsample_masks_matches_interpolate():
    sd = synth.make_params(seed=2, num_queries=8)
    head = build_head_from_state_dict(sd, 8, 49, 'fp32', DEV)
    g = torch.Generator().manual_seed(4)
    for dtype in (torch.float32, torch.bfloat16):
        x = torch.randn((2, 8, 66, 50), generator=g).to(DEV).to(dtype)
        want = F.interpolate(x.float(), size=(264, 200), mode='bilinear', align_corners=False)
        got = P.upsample_masks(head, x, (264, 200))
        assert float((got - want).abs().max()) < 1e-5


class _StubPixelDecoder(torch.nn.Module):
    """Stands in for mmdet's MSDeformAttnPixelDecoder (the step BEFORE the path): passes precomputed features through."""

    def forward(self, feats):
        return feats[0], list(feats[1:])


def _caption_head(Q, precision='fp32'):
    from cgg_b200.head import Mask2FormerHeadOpenB200
    sd = synth.make_params(seed=6, num_queries=Q, perturb=True)
    head = Mask2FormerHeadOpenB200(num_things_classes=48, num_stuff_classes=0, num_queries=Q, precision=precision,
                                   use_class_emb=True, use_caption=True, bert_vocab_size=1500,
                                   loss_grounding=dict(type='GroundingLoss', loss_weight=2.0),
                                   pixel_decoder=_StubPixelDecoder())
    ids, cap_mask, table, lw, lb = synth.make_captions(6, 2, vocab=1500)
    ids[0, :4] = torch.tensor([1001, 1100, 1250, 1499])          # image 0: four nouns; image 1: an empty caption
    cap_mask[0, :4] = 1
    ids[0, 4:] = 0
    cap_mask[0, 4:] = 0
    sd = dict(sd)
    sd['bert_embeddings.word_embeddings.weight'] = table
    sd['bert_embeddings.LayerNorm.weight'] = lw + 0.1
    sd['bert_embeddings.LayerNorm.bias'] = lb - 0.05
    head.load_state_dict(sd, strict=True)
    return head.to(DEV), sd, ids, cap_mask


def test_simple_test_contract_and_att():
    """mask2former_head.py:923-980: (assigned_labels, cls_emb[-1], upsampled mask[-1], None, att)."""
    from oracle import cgg_oracle as O
    Q, B = 24, 2
    head, sd, ids, cap_mask = _caption_head(Q)
    head.eval()
    mf, mems = synth.make_inputs(8, B, 128, 160)
    feats = [mf.to(DEV)] + [m.to(DEV) for m in mems]
    metas = [dict(batch_input_shape=(128, 160), img_shape=(128, 160, 3), ori_shape=(128, 160, 3)) for _ in range(B)]
    noun_ids = ids[0, :5].to(DEV)
    with torch.no_grad():
        labels, emb, masks, cap, att = head.simple_test(feats, metas, with_att=True, nouns_ids=noun_ids)
    ref = O.decoder_forward({k: v for k, v in sd.items() if not k.startswith('bert')}, mf, mems)
    want_mask = F.interpolate(ref['mask'][9], size=(128, 160), mode='bilinear', align_corners=False)
    assert cap is None and tuple(masks.shape) == (B, Q, 128, 160)
    assert float((masks.cpu() - want_mask).abs().max()) < 2e-3 * float(want_mask.abs().max())
    assert float((labels.cpu() - ref['cls'][9]).abs().max()) < 2e-3 * float(ref['cls'][9].abs().max())
    nouns = O.noun_embeddings(sd['bert_embeddings.word_embeddings.weight'], sd['bert_embeddings.LayerNorm.weight'],
                              sd['bert_embeddings.LayerNorm.bias'], ids[0, :5])
    want_att = ref['emb'][9][0] @ nouns.t()
    assert float((att.cpu() - want_att).abs().max()) < 2e-3 * float(want_att.abs().max())


def test_forward_train_grounding_losses_and_gradients():
    """mask2former_head.py:851-921 -> loss :393-462 for the on-path term: 10 grounding losses under the reference's
    key names, weight 2.0, autograd-connected to the head's parameters."""
    from oracle import cgg_oracle as O
    Q, B = 24, 2
    head, sd, ids, cap_mask = _caption_head(Q)
    head.train()
    mf, mems = synth.make_inputs(8, B, 128, 160)
    feats = [mf.to(DEV)] + [m.to(DEV) for m in mems]
    metas = [dict() for _ in range(B)]
    losses = head.forward_train(feats, metas, None, None, None, None, None, None, list(ids.to(DEV)), list(cap_mask.to(DEV)))
    assert set(losses) == {'loss_grounding'} | {'d%d.loss_grounding' % j for j in range(9)}
    sd_o = {k: v for k, v in sd.items() if not k.startswith('bert')}
    ref = O.decoder_forward(sd_o, mf, mems)
    nouns = O.noun_embeddings(sd['bert_embeddings.word_embeddings.weight'], sd['bert_embeddings.LayerNorm.weight'],
                              sd['bert_embeddings.LayerNorm.bias'], ids)
    for j in range(10):
        want = float(O.grounding_loss(ref['emb'][j], nouns, cap_mask, 10.0, 2.0))
        got = float(losses['loss_grounding' if j == 9 else 'd%d.loss_grounding' % j])
        assert abs(got - want) < 2e-3 * max(1.0, abs(want)), (j, got, want)
    sum(losses.values()).backward()
    assert head.v2l_transform.weight.grad is not None and float(head.v2l_transform.weight.grad.abs().max()) > 0
    assert head.query_feat.weight.grad is not None and head.bert_embeddings.word_embeddings.weight.grad is None
