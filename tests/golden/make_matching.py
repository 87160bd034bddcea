"""Generates tests/golden/matching.npz: inputs and outputs of `loss_single` (mask2former_head.py:464-629) of the UNMODIFIED
reference head (run in this container through oracle/ref_shim.py) for the matching-based losses -- the fixture that
travels to the GPU box, where /root/reference does not exist.
    python tests/golden/make_matching.py"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim                 # noqa: E402
from test_matching_cpu import make_case     # noqa: E402


def main():
    ncls, P, seed = 48, 1024, 7
    R = ref_shim.REF_ROOT
    head = ref_shim.build_reference_head(with_losses=True, num_points=P, num_queries=20, num_known=ncls,
                                         known_file=R + '/datasets/unknown/known_65.txt',
                                         unknown_file=R + '/datasets/unknown/unknown_17.txt')
    head.use_caption = head.use_caption_generation = head.use_caption_align = False
    head._get_cls_emb_logits = lambda preds: preds
    cls_scores, cls_emb_logits, mask_preds, gt_labels, gt_masks = make_case(seed, B=3, gts=(4, 0, 6))
    metas = [dict() for _ in gt_labels]
    torch.manual_seed(seed)
    loss_cls, loss_cls_emb, _, _, _, loss_mask, loss_dice = head.loss_single(
        cls_scores, cls_emb_logits, mask_preds, gt_labels, gt_masks, None, None, None, None, None, None, metas)
    torch.manual_seed(seed)                 # the assignment alone: the same first torch.rand draws
    labels_list = head.get_targets(list(cls_scores), list(cls_emb_logits), list(mask_preds), gt_labels, gt_masks, metas)[0]
    out = dict(B=3, ncls=ncls, num_points=P, seed=seed, cls_scores=cls_scores.numpy(), cls_emb_logits=cls_emb_logits.numpy(),
               mask_preds=mask_preds.numpy(), labels=torch.stack(labels_list, 0).numpy(),
               loss_cls=float(loss_cls), loss_cls_emb=float(loss_cls_emb), loss_mask=float(loss_mask), loss_dice=float(loss_dice))
    for b in range(3):
        out['gt_labels_%d' % b] = gt_labels[b].numpy()
        out['gt_masks_%d' % b] = gt_masks[b].numpy()
    np.savez_compressed(os.path.join(HERE, 'matching.npz'), **out)
    print({k: v for k, v in out.items() if k.startswith('loss')})


if __name__ == '__main__':
    main()
