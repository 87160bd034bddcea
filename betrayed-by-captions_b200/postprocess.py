"""Test-time step after the decoder path: the final mask upsample of
`Mask2FormerHeadOpen.simple_test` (open_set/models/mask2former_head.py:957-964) and the instance scoring of
`MaskFormerFusionHeadOpen` (open_set/models/maskformer_fusion_head.py:297-366, :412-425), on the kernels behind
`cgg_upsample_masks`, `cgg_instance_mask_stats`, `cgg_similarity` and `cgg_softmax_rows`.

`instance_postprocess_emb_fused` returns what the reference's `instance_postprocess_emb` returns for every image of the
batch -- labels, boxes with detection scores, binary masks -- but computes the masks' `> 0` bits, pixel counts, sigmoid
sums and bounding boxes in ONE pass over the (B, Q, H/4, W/4) logits: the (B, Q, H, W) fp32 logits that the reference
materialises (6.7 GB for 16 images at 1024^2) never exist.  Masks come back bit-packed (32 pixels per word);
`unpack_masks` expands them when a dense bool tensor is wanted."""
import ctypes as C

import torch

from . import lib as _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _ctx(head, device):
    rt = head._runtime(device)
    return rt, rt.lib, rt.handle, C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def upsample_masks(head, mask_pred, size):
    """F.interpolate(mask_pred, size, mode='bilinear', align_corners=False) -> fp32 (head.py:957-964)."""
    if not mask_pred.is_cuda:
        raise _lib.CggError('upsample_masks runs on CUDA only')
    rt, lib, h, s = _ctx(head, mask_pred.device)
    mp = mask_pred.contiguous()
    B, Q, h4, w4 = mp.shape
    out = torch.empty((B, Q, size[0], size[1]), dtype=torch.float32, device=mp.device)
    assert mp.dtype in (torch.float32, torch.bfloat16)
    _lib.check(lib.cgg_upsample_masks(h, _p(mp), int(mp.dtype == torch.bfloat16), _p(out), B * Q, h4, w4, size[0], size[1], s),
               h, 'cgg_upsample_masks')
    return out


def instance_mask_stats(head, mask_pred, up_size, crops, outs=None, want_bits=True):
    """mask_pred (B, Q, h4, w4) last-call logits; up_size = batch_input_shape; crops[b] = img_shape (h, w); outs[b] =
    ori_shape (h, w) when rescaling (None: no rescale).  Returns dict(bits (B,Q,Hmax,W32) int32 | None, count (B,Q) int32,
    sig_sum (B,Q) fp32, bbox (B,Q,4) int32, out_sizes)."""
    rt, lib, h, s = _ctx(head, mask_pred.device)
    mp = mask_pred.contiguous()
    B, Q, h4, w4 = mp.shape
    outs = outs if outs is not None else crops
    geom = torch.tensor([[c[0], c[1], o[0], o[1]] for c, o in zip(crops, outs)], dtype=torch.int32).to(mp.device)
    Hm, Wm = max(o[0] for o in outs), max(o[1] for o in outs)
    dev = mp.device
    bits = torch.zeros((B, Q, Hm, (Wm + 31) // 32), dtype=torch.int32, device=dev) if want_bits else None
    count = torch.empty((B, Q), dtype=torch.int32, device=dev)
    sig = torch.empty((B, Q), dtype=torch.float32, device=dev)
    bbox = torch.empty((B, Q, 4), dtype=torch.int32, device=dev)
    _lib.check(lib.cgg_instance_mask_stats(h, _p(mp), int(mp.dtype == torch.bfloat16), _p(geom), B, Q, h4, w4, up_size[0],
                                           up_size[1], Hm, Wm, _p(bits), _p(count), _p(sig), _p(bbox), s),
               h, 'cgg_instance_mask_stats')
    return dict(bits=bits, count=count, sig_sum=sig, bbox=bbox, out_sizes=list(outs))


def unpack_masks(bits, width):
    """(..., H, W32) int32 words -> (..., H, width) bool."""
    sh = torch.arange(32, device=bits.device, dtype=torch.int32)
    b = ((bits.unsqueeze(-1) >> sh) & 1).bool()
    return b.flatten(-2)[..., :width]


def cls_emb_scores(head, cls_emb_preds, class_embs):
    """get_cls_emb_scores (maskformer_fusion_head.py:297-315): softmax(emb @ class_embs^T) over the classes."""
    rt, lib, h, s = _ctx(head, cls_emb_preds.device)
    x = cls_emb_preds.reshape(-1, cls_emb_preds.shape[-1])
    scores = rt.similarity(x, class_embs, 1.0)
    _lib.check(lib.cgg_softmax_rows(h, _p(scores), scores.shape[0], scores.shape[1], s), h, 'cgg_softmax_rows')
    return scores.view(*cls_emb_preds.shape[:-1], class_embs.shape[0])


def instance_postprocess_emb_fused(head, mask_cls_emb, mask_pred, class_embs, img_metas, rescale=False, max_per_image=100):
    """`instance_postprocess_emb` (maskformer_fusion_head.py:318-366) for the whole batch, from the LOW-resolution logits.
    mask_cls_emb (B, Q, d_l); mask_pred (B, Q, h4, w4).  Per image: (labels (n,), bboxes (n, 5), packed masks (n, H, W32),
    (H, W))."""
    B, Q = mask_pred.shape[:2]
    up = img_metas[0]['batch_input_shape']
    crops = [m['img_shape'][:2] for m in img_metas]
    outs = [m['ori_shape'][:2] for m in img_metas] if rescale else None
    st = instance_mask_stats(head, mask_pred, up, crops, outs)
    scores_all = cls_emb_scores(head, mask_cls_emb, class_embs)[..., :-1]            # drop the void column
    ncls = scores_all.shape[-1]
    results = []
    for b in range(B):
        sc, top = scores_all[b].flatten().topk(max_per_image, sorted=False)
        labels = top % ncls
        qi = top // ncls
        cnt = st['count'][b, qi].float()
        mask_score = st['sig_sum'][b, qi] / (cnt + 1e-6)
        det = sc * mask_score
        boxes = torch.cat([st['bbox'][b, qi].float(), det[:, None]], dim=-1)
        results.append((labels, boxes, st['bits'][b, qi][:, :st['out_sizes'][b][0]], st['out_sizes'][b]))
    return results
