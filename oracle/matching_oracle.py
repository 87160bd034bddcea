"""TEST INFRASTRUCTURE -- CPU restatement of the step after the path at training time: the
point-sampled mask / dice losses, the class losses and the matching cost matrices of `loss_single`
(open_set/models/mask2former_head.py:464-629) with its target assignment (`_get_target_single`, :320-390; assigner
open_set/assigners/mask_hungarian_assigner.py:47-146).  Only tests/, __graft_entry__.smoke() and bench.py's baseline legs
may import this file; the product (betrayed-by-captions_b200/matching.py) never does.

The reference calls into two third-party packages that are NOT in /root/reference: mmcv-full 1.7.1 and mmdet 2.28.2
(pinned in the reference's README.md:54,63).  Their published algorithms are restated here in plain torch, each function
naming the upstream file it follows:
  * mmcv/ops/point_sample.py            point_sample (grid_sample on 2*p - 1, bilinear, zeros padding, align_corners=False)
  * mmdet/models/utils/point_sample.py  get_uncertainty, get_uncertain_point_coords_with_randomness
  * mmdet/core/bbox/match_costs/match_cost.py   ClassificationCost, CrossEntropyLossCost (use_sigmoid), DiceCost
  * mmdet/core/bbox/samplers/mask_pseudo_sampler.py   MaskPseudoSampler (positives = assigned, negatives = the rest)
  * mmdet/models/losses/dice_loss.py    dice_loss (naive_dice), mmdet/models/losses/utils.py weight_reduce_loss
  * mmdet/models/losses/cross_entropy_loss.py   cross_entropy / binary_cross_entropy -- the reference carries its own copy
    of these two (open_set/models/losses/cross_entropy_loss.py:62-110, :138-199), which IS in /root/reference.
Pinned by tests/test_matching_cpu.py: the whole of `loss_single` of the UNMODIFIED reference head (imported through
oracle/ref_shim.py, which plugs these same restatements in for the absent mmcv / mmdet entry points and registers the
reference's own CrossEntropyLossOpen for the config's 'CrossEntropyLoss') against `loss_single_matching` below on the
same seeded inputs and the same torch.rand stream, and by the committed fixture tests/golden/matching.npz.
"""
import torch
import torch.nn.functional as F

try:
    from scipy.optimize import linear_sum_assignment
except ImportError:                                     # pragma: no cover
    linear_sum_assignment = None

EPS32 = torch.finfo(torch.float32).eps


# ------------------------------------------------------------------------------------------- mmcv / mmdet utilities
def point_sample(inp, points, align_corners=False):
    """mmcv/ops/point_sample.py: inp (N, C, H, W), points (N, P, 2) in [0, 1] x [0, 1] as (x, y) -> (N, C, P)."""
    out = F.grid_sample(inp, (2.0 * points - 1.0).unsqueeze(2), align_corners=align_corners)
    return out.squeeze(3)


def get_uncertain_point_coords_with_randomness(mask_pred, num_points, oversample_ratio, importance_sample_ratio):
    """mmdet/models/utils/point_sample.py (labels=None, one channel): oversample uniformly, keep the
    importance_sample_ratio * num_points most uncertain points (-|logit| largest), fill up with fresh uniform points.
    Consumes torch.rand twice, in this order: (N, int(num_points * oversample_ratio), 2) then (N, num_random, 2)."""
    assert oversample_ratio >= 1 and 0 <= importance_sample_ratio <= 1
    n = mask_pred.shape[0]
    num_sampled = int(num_points * oversample_ratio)
    coords = torch.rand(n, num_sampled, 2, device=mask_pred.device)
    logits = point_sample(mask_pred, coords)
    unc = -torch.abs(logits)                                            # get_uncertainty, single channel
    num_uncertain = int(importance_sample_ratio * num_points)
    num_random = num_points - num_uncertain
    idx = torch.topk(unc[:, 0, :], k=num_uncertain, dim=1)[1]
    shift = num_sampled * torch.arange(n, dtype=torch.long, device=mask_pred.device)
    idx = idx + shift[:, None]
    coords = coords.view(-1, 2)[idx.view(-1), :].view(n, num_uncertain, 2)
    if num_random > 0:
        coords = torch.cat((coords, torch.rand(n, num_random, 2, device=mask_pred.device)), dim=1)
    return coords


def weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """mmdet/models/losses/utils.py."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return loss.mean() if reduction == 'mean' else (loss.sum() if reduction == 'sum' else loss)
    if reduction == 'mean':
        return loss.sum() / (avg_factor + EPS32)
    if reduction == 'none':
        return loss
    raise ValueError('avg_factor can not be used with reduction="sum"')


def cross_entropy_loss(pred, label, weight, avg_factor, class_weight, loss_weight):
    """CrossEntropyLoss(use_sigmoid=False, reduction='mean', class_weight=...): cross_entropy_loss.py:62-110 (reference copy)."""
    loss = F.cross_entropy(pred, label, weight=class_weight, reduction='none')
    return loss_weight * weight_reduce_loss(loss, weight.float() if weight is not None else None, 'mean', avg_factor)


def binary_cross_entropy_loss(pred, label, avg_factor, loss_weight):
    """CrossEntropyLoss(use_sigmoid=True, reduction='mean') on (N,) logits / targets: cross_entropy_loss.py:138-199."""
    valid = ((label >= 0) & (label != -100)).float()
    loss = F.binary_cross_entropy_with_logits(pred, label.float(), reduction='none')
    return loss_weight * weight_reduce_loss(loss, valid, 'mean', avg_factor)


def dice_loss(pred, target, avg_factor, loss_weight, eps=1.0):
    """DiceLoss(use_sigmoid=True, activate=True, reduction='mean', naive_dice=True, eps): mmdet/models/losses/dice_loss.py."""
    inp = pred.sigmoid().flatten(1)
    target = target.flatten(1).float()
    a = torch.sum(inp * target, 1)
    b = torch.sum(inp, 1)
    c = torch.sum(target, 1)
    loss = 1 - (2 * a + eps) / (b + c + eps)
    return loss_weight * weight_reduce_loss(loss, None, 'mean', avg_factor)


# ---------------------------------------------------------------------------------------------------- match costs
def classification_cost(cls_pred, gt_labels, weight):
    """ClassificationCost: -softmax(cls_pred)[:, gt_labels] * weight."""
    return -cls_pred.softmax(-1)[:, gt_labels] * weight


def cross_entropy_loss_cost(mask_pred, gt_mask, weight):
    """CrossEntropyLossCost(use_sigmoid=True): (BCE(x, 1) . g + BCE(x, 0) . (1 - g)) / n over the sampled points."""
    x = mask_pred.flatten(1).float()
    g = gt_mask.flatten(1).float()
    n = x.shape[1]
    pos = F.binary_cross_entropy_with_logits(x, torch.ones_like(x), reduction='none')
    neg = F.binary_cross_entropy_with_logits(x, torch.zeros_like(x), reduction='none')
    cost = torch.einsum('nc,mc->nm', pos, g) + torch.einsum('nc,mc->nm', neg, 1 - g)
    return cost / n * weight


def dice_cost(mask_pred, gt_mask, weight, pred_act=True, eps=1.0, naive_dice=True):
    """DiceCost: 1 - (2 p.g + eps) / (sum p + sum g + eps), p = sigmoid(x) when pred_act."""
    p = mask_pred.sigmoid() if pred_act else mask_pred
    p = p.flatten(1)
    g = gt_mask.flatten(1).float()
    num = 2 * torch.einsum('nc,mc->nm', p, g)
    if naive_dice:
        den = p.sum(-1)[:, None] + g.sum(-1)[None, :]
    else:
        den = p.pow(2).sum(1)[:, None] + g.pow(2).sum(1)[None, :]
    return (1 - (num + eps) / (den + eps)) * weight


def matching_cost(cls_pred, cls_emb_logit, mask_points_pred, gt_labels, gt_points_masks, w):
    """mask_hungarian_assigner.py:98-125: cost = cls + cls_emb + mask + dice (a term with weight 0 is skipped)."""
    cost = 0
    if w['cls'] != 0 and cls_pred is not None:
        cost = cost + classification_cost(cls_pred, gt_labels, w['cls'])
    if w['cls_emb'] != 0 and cls_emb_logit is not None:
        cost = cost + classification_cost(cls_emb_logit, gt_labels, w['cls_emb'])
    if w['mask'] != 0:
        cost = cost + cross_entropy_loss_cost(mask_points_pred, gt_points_masks, w['mask'])
    if w['dice'] != 0:
        cost = cost + dice_cost(mask_points_pred, gt_points_masks, w['dice'], eps=w.get('dice_eps', 1.0))
    return cost


def assign(cost, num_query, num_gt):
    """mask_hungarian_assigner.py:84-146 + MaskPseudoSampler: (pos_inds, pos_assigned_gt_inds) of the Hungarian matching."""
    if num_gt == 0 or num_query == 0:
        e = torch.zeros((0,), dtype=torch.long)
        return e, e
    rows, cols = linear_sum_assignment(cost.detach().cpu())
    assigned = torch.zeros((num_query,), dtype=torch.long)
    assigned[torch.from_numpy(rows)] = torch.from_numpy(cols) + 1
    pos = torch.nonzero(assigned > 0, as_tuple=False).squeeze(-1).unique()
    return pos, assigned[pos] - 1


DEFAULT_CFG = dict(num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75,
                   cost=dict(cls=0.0, cls_emb=2.0, mask=5.0, dice=5.0, dice_eps=1.0),
                   loss_cls_weight=0.0, loss_cls_emb_weight=2.0, loss_mask_weight=5.0, loss_dice_weight=5.0, dice_eps=1.0,
                   bg_class_weight=0.1)     # configs/openset_panoptic/coco_panoptic_p20.py:111-139, :163-175


def get_target_single(cls_score, cls_emb_logit, mask_pred, gt_labels, gt_masks, num_classes, cfg):
    """mask2former_head.py:320-390.  Consumes torch.rand((1, num_points, 2)) once."""
    num_queries, num_gts = cls_score.shape[0], gt_labels.shape[0]
    point_coords = torch.rand((1, cfg['num_points'], 2), device=cls_score.device)
    mask_points_pred = point_sample(mask_pred.unsqueeze(1), point_coords.repeat(num_queries, 1, 1)).squeeze(1)
    gt_points_masks = point_sample(gt_masks.unsqueeze(1).float(), point_coords.repeat(num_gts, 1, 1)).squeeze(1) \
        if num_gts > 0 else gt_masks.new_zeros((0, cfg['num_points']), dtype=torch.float32)
    cost = matching_cost(cls_score, cls_emb_logit, mask_points_pred, gt_labels, gt_points_masks, cfg['cost']) \
        if num_gts > 0 else None
    pos_inds, pos_gt = assign(cost, num_queries, num_gts)
    labels = gt_labels.new_full((num_queries,), num_classes, dtype=torch.long)
    labels[pos_inds] = gt_labels[pos_gt]
    mask_weights = mask_pred.new_zeros((num_queries,))
    mask_weights[pos_inds] = 1.0
    return labels, gt_masks[pos_gt], mask_weights, pos_inds, cost


def loss_single_matching(cls_scores, cls_emb_logits, mask_preds, gt_labels_list, gt_masks_list, num_classes, cfg=None):
    """The matching-based terms of loss_single (mask2former_head.py:464-629): returns dict(loss_cls, loss_cls_emb,
    loss_mask, loss_dice) plus the assignment (labels (B, Q), mask_weights (B, Q)) for inspection.
    cls_scores (B, Q, C+1), cls_emb_logits (B, Q, C+1) (already `_get_cls_emb_logits`, :631-648) or None,
    mask_preds (B, Q, h, w); gt_labels_list / gt_masks_list per image ((G,), (G, h, w))."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    B = cls_scores.shape[0]
    labels, mask_targets, mask_weights, num_pos = [], [], [], 0
    for i in range(B):
        lab, mt, mw, pos, _ = get_target_single(cls_scores[i], None if cls_emb_logits is None else cls_emb_logits[i],
                                                mask_preds[i], gt_labels_list[i], gt_masks_list[i], num_classes, cfg)
        labels.append(lab), mask_targets.append(mt), mask_weights.append(mw)
        num_pos += pos.numel()
    labels = torch.stack(labels, 0)
    mask_targets = torch.cat(mask_targets, 0)
    mask_weights = torch.stack(mask_weights, 0)
    class_weight = cls_scores.new_tensor([1.0] * num_classes + [cfg['bg_class_weight']])
    flat_labels = labels.flatten(0, 1)
    label_weights = torch.ones_like(flat_labels)
    avg = class_weight[flat_labels].sum()
    out = dict(labels=labels, mask_weights=mask_weights)
    out['loss_cls'] = cross_entropy_loss(cls_scores.flatten(0, 1), flat_labels, label_weights, avg, class_weight,
                                         cfg['loss_cls_weight'])
    out['loss_cls_emb'] = cls_scores.new_tensor(0.0)
    if cls_emb_logits is not None:
        out['loss_cls_emb'] = cross_entropy_loss(cls_emb_logits.flatten(0, 1), flat_labels, label_weights.float(), avg,
                                                 class_weight, cfg['loss_cls_emb_weight'])
    num_total_masks = max(float(num_pos), 1.0)                                  # reduce_mean over ranks: one rank here
    pos_preds = mask_preds[mask_weights > 0]
    if mask_targets.shape[0] == 0:                                              # zero match (:582-586)
        out['loss_dice'] = pos_preds.sum()
        out['loss_mask'] = pos_preds.sum()
        return out
    with torch.no_grad():
        coords = get_uncertain_point_coords_with_randomness(pos_preds.unsqueeze(1), cfg['num_points'],
                                                            cfg['oversample_ratio'], cfg['importance_sample_ratio'])
        point_targets = point_sample(mask_targets.unsqueeze(1).float(), coords).squeeze(1)
    point_preds = point_sample(pos_preds.unsqueeze(1), coords).squeeze(1)
    out['loss_dice'] = dice_loss(point_preds, point_targets, num_total_masks, cfg['loss_dice_weight'], cfg['dice_eps'])
    out['loss_mask'] = binary_cross_entropy_loss(point_preds.reshape(-1), point_targets.reshape(-1),
                                                 num_total_masks * cfg['num_points'], cfg['loss_mask_weight'])
    out['point_coords'] = coords
    return out
