"""Training-side grounding loss of the path.

`GroundingLossB200` mirrors `open_set/models/losses/grounding_loss.py:79-91` (module taking
(cls_emb_pred, gt_caption_embs, gt_caption_mask, temperature), times `loss_weight`); the value
and the gradient w.r.t. the predictions come from the CUDA kernels behind `cgg_grounding_loss`
and `cgg_grounding_loss_backward`.  Captions are frozen BERT rows in the reference
(`mask2former_head.py:251-254`), so no caption gradient exists.

`gather_captions_and_preds` replaces `mask2former_head.py:650-684`: the reference issues three
all_gathers per head call (30 per step); captions are identical for the 10 head calls, so here
they travel once and the predictions of all head calls travel stacked -- two collectives per
step, same results, gradients only into the local slot (`:678`).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib as _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class _Handle:
    """Process-wide C-ABI handle for the weight-free grounding entry points, one per device."""
    _handles = {}

    @classmethod
    def get(cls, device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key not in cls._handles:
            lib = _lib.load()
            h = C.c_void_p()
            cfg = _lib.Config(100, 256, 8, 2048, 9, 49, 768, _lib.FP32, 0)
            with torch.cuda.device(key[1]):
                _lib.check(lib.cgg_create(C.byref(h), C.byref(cfg)), None, 'cgg_create')
            cls._handles[key] = (lib, h)
        return cls._handles[key]


class _GroundingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, cap, cap_mask, temperature, loss_weight):
        if not pred.is_cuda:
            raise RuntimeError('GroundingLossB200 runs on a CUDA device only (no CPU path)')
        lib, h = _Handle.get(pred.device)
        p = pred.detach().float().contiguous()
        c = cap.detach().float().contiguous()
        m = cap_mask.to(torch.int64).contiguous()
        Bg, Q, D = p.shape
        T = c.shape[1]
        stream = C.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)
        nbytes = lib.cgg_grounding_scratch_bytes(Bg, Q, T)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=pred.device)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        st = lib.cgg_grounding_loss(h, _ptr(p), _ptr(c), _ptr(m), Bg, Q, T, D, float(temperature), float(loss_weight),
                                    _ptr(loss), _ptr(scratch), nbytes, stream)
        _lib.check(st, h, 'cgg_grounding_loss')
        ctx.save_for_backward(p, c, m)
        ctx.temperature, ctx.loss_weight, ctx.in_dtype = float(temperature), float(loss_weight), pred.dtype
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        p, c, m = ctx.saved_tensors
        lib, h = _Handle.get(p.device)
        Bg, Q, D = p.shape
        T = c.shape[1]
        stream = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        nbytes = lib.cgg_grounding_bwd_scratch_bytes(Bg, Q, T)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
        dpred = torch.empty_like(p)
        # the upstream gradient is a device scalar; the kernels take it pre-multiplied into d(cost), so
        # compute with 1.0 and scale the result on the device (no host sync)
        st = lib.cgg_grounding_loss_backward(h, _ptr(p), _ptr(c), _ptr(m), Bg, Q, T, D, ctx.temperature,
                                             ctx.loss_weight, 1.0, _ptr(dpred), _ptr(scratch), nbytes, stream)
        _lib.check(st, h, 'cgg_grounding_loss_backward')
        return (dpred * grad_out).to(ctx.in_dtype), None, None, None, None


def grounding_loss(cls_emb_pred, gt_caption_embs, gt_caption_mask, temperature, loss_weight=1.0):
    """`losses/grounding_loss.py:9-77` with autograd support (gradient w.r.t. cls_emb_pred)."""
    return _GroundingFn.apply(cls_emb_pred, gt_caption_embs, gt_caption_mask, temperature, loss_weight)


class GroundingLossB200(torch.nn.Module):
    """Drop-in for `GroundingLoss` (`losses/grounding_loss.py:79-91`)."""

    def __init__(self, loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, cls_emb_pred, gt_caption_embs, gt_caption_mask, temperature):
        return grounding_loss(cls_emb_pred, gt_caption_embs, gt_caption_mask, temperature, self.loss_weight)


class _SimilarityFn(torch.autograd.Function):
    """out = scale * a @ b^T through `cgg_similarity`; both gradients are the same contraction on transposed
    operands (da = scale * dout @ b, db = scale * dout^T @ a), so the backward uses the same entry point."""

    @staticmethod
    def _sim(a, b, scale):
        lib, h = _Handle.get(a.device)
        a = a.detach().float().contiguous()
        b = b.detach().float().contiguous()
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
        stream = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
        st = lib.cgg_similarity(h, _ptr(a), _ptr(b), a.shape[0], b.shape[0], a.shape[1], float(scale), _ptr(out), stream)
        _lib.check(st, h, 'cgg_similarity')
        return out

    @staticmethod
    def forward(ctx, a, b, scale):
        if not a.is_cuda:
            raise RuntimeError('similarity runs on a CUDA device only (no CPU path)')
        ctx.save_for_backward(a, b)
        ctx.scale = float(scale)
        return _SimilarityFn._sim(a, b, scale)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = db = None
        if ctx.needs_input_grad[0]:
            da = _SimilarityFn._sim(g, b.t(), ctx.scale).to(a.dtype)          # (M,N) x (D,N)^T -> (M,D)
        if ctx.needs_input_grad[1]:
            db = _SimilarityFn._sim(g.t(), a.t(), ctx.scale).to(b.dtype)      # (N,M) x (D,M)^T -> (N,D)
        return da, db, None


def similarity(a, b, scale=1.0):
    """scale * a @ b^T (`_get_cls_emb_logits`, head.py:631-648; test-time `att`, :973-978), differentiable."""
    return _SimilarityFn.apply(a, b, scale)


class _GatherKeepLocalGrad(torch.autograd.Function):
    """all_gather along dim `dim`; backward hands the local slot's gradient back (the other slots are
    detached copies, exactly like the re-insertion at `mask2former_head.py:678`)."""

    @staticmethod
    def forward(ctx, x, dim, group):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        x = x.contiguous()
        out = torch.empty((world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        if x.is_cuda:
            dist.all_gather_into_tensor(out, x, group=group)
        else:                                   # gloo (CPU tests of the host logic)
            dist.all_gather(list(out.unbind(0)), x, group=group)
        ctx.dim, ctx.rank, ctx.n = dim, rank, x.shape[dim]
        # (world, ..., n, ...) -> (..., world*n, ...)
        out = out.movedim(0, dim)
        shape = list(x.shape)
        shape[dim] = world * x.shape[dim]
        return out.reshape(shape)

    @staticmethod
    def backward(ctx, g):
        return g.narrow(ctx.dim, ctx.rank * ctx.n, ctx.n).contiguous(), None, None


def gather_captions_and_preds(gt_caption_embs_list, gt_caption_mask_list, cls_emb_preds, group=None):
    """`mask2former_head.py:650-684`, batched.

    gt_caption_embs_list / gt_caption_mask_list: per-image (T, D) / (T,) tensors (or already stacked).
    cls_emb_preds: (B, Q, D) for one head call, or (L, B, Q, D) for all head calls stacked.
    Returns (all_embs (B*world, T, D), all_mask (B*world, T), all_preds (B*world, Q, D) or (L, B*world, Q, D)),
    rank-major along the batch axis like the reference's torch.cat of the gathered lists.
    """
    embs = torch.stack(list(gt_caption_embs_list), 0) if not torch.is_tensor(gt_caption_embs_list) \
        else gt_caption_embs_list
    mask = torch.stack(list(gt_caption_mask_list), 0) if not torch.is_tensor(gt_caption_mask_list) \
        else gt_caption_mask_list
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return embs, mask, cls_emb_preds
    B, T, D = embs.shape
    # captions: one collective for embeddings + mask (the mask rides as one extra fp32 column block; 0/1 is exact)
    packed = torch.cat([embs.float().reshape(B, T * D), mask.to(torch.float32)], dim=1)
    allp = _GatherKeepLocalGrad.apply(packed.detach(), 0, group)
    all_embs = allp[:, :T * D].reshape(-1, T, D).to(embs.dtype)
    all_mask = allp[:, T * D:].round().to(mask.dtype)
    batch_dim = cls_emb_preds.dim() - 3
    all_preds = _GatherKeepLocalGrad.apply(cls_emb_preds, batch_dim, group)
    return all_embs, all_mask, all_preds
