"""The caption generator of the open-vocabulary head: `CaptionTransformer`
(open_set/models/transformers/caption_tranformer.py:20-43: adapter, sinusoidal positions, 4 post-norm decoder blocks of
masked self-attention / cross-attention over the query embeddings / FFN, open_set/models/transformers/transformers.py
:58-134, :180-234, :252-267, and the 30 522-way generator), its training loss (`loss_caption_generation`,
open_set/models/mask2former_head.py:552-583) and the test-time beam search (open_set/utils/eval/inference.py:84-157).

The module carries the reference's state_dict keys (checkpoints load unchanged); its forward AND backward are the stage
kernels of the C-ABI library through the autograd nodes of train.py -- every linear layer a `cgg_gemm_f32` call (tcgen05
kind::tf32 in the tf32 training mode, fp32 FMA otherwise), the attention as the (image, head)-batched products with the
row-softmax kernels (`_AttentionViews`: 8 heads x 96, causal + key-padding masks as a bitmap), LayerNorm, broadcast adds.
The beam search keeps the reference's control flow (host-side candidate bookkeeping, top-k over the flattened beams) and
runs the network in fp32 FMA mode, so that the token sequences are the reference's."""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib as _lib
from .train import _K, _Linear, _LayerNorm, _AddRows, _AttentionViews

BOS_TOKEN, EOS_TOKEN = 101, 102             # bert-base-uncased [CLS] / [SEP] (mask2former_head.py:30-31)


class _SelfAttn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv_layer = nn.Linear(dim, 3 * dim)
        self.out_layer = nn.Linear(dim, dim)


class _CrossAttn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.to_qry, self.to_key, self.to_val, self.to_out = (nn.Linear(dim, dim) for _ in range(4))


class _FFN(nn.Module):
    """FeedForwardNetwork([in, ff, in], [ReLU, Identity], [drop, 0]): keys linears.{0,1}.0.*"""

    def __init__(self, dim, ff, drop):
        super().__init__()
        self.linears = nn.ModuleList([
            nn.Sequential(nn.Linear(dim, ff), nn.Dropout(drop) if drop > 0.0 else nn.Identity(), nn.ReLU()),
            nn.Sequential(nn.Linear(ff, dim), nn.Identity(), nn.Identity())])


class _Block(nn.Module):
    def __init__(self, dim, ff, heads, drop, pre_norm):
        super().__init__()
        self.mha_layer, self.crx_layer, self.ffn_layer = _SelfAttn(dim), _CrossAttn(dim), _FFN(dim, ff, drop)
        self.dropout_layer = nn.ModuleDict({k: nn.Dropout(drop) for k in ('mha', 'crx', 'ffn')})
        self.layer_normalz = nn.ModuleDict({
            k: nn.ModuleList([nn.LayerNorm(dim) if pre_norm else nn.Identity(),
                              nn.LayerNorm(dim) if not pre_norm else nn.Identity()]) for k in ('mha', 'crx', 'ffn')})


class _Decoder(nn.Module):
    def __init__(self, n, dim, ff, heads, drop, pre_norm):
        super().__init__()
        self.decoders = nn.ModuleList([_Block(dim, ff, heads, drop, pre_norm) for _ in range(n)])


class _Positions(nn.Module):
    """PositionalEncoding (transformers.py:9-25): sin on even, cos on odd channels of pos / 10000^((j - j%2)/dim)."""

    def __init__(self, seq_length, dim, drop):
        super().__init__()
        pos = np.arange(0, seq_length)[:, None]
        idx = np.fromfunction(lambda _, j: j - j % 2, shape=(1, dim))
        mask = np.fromfunction(lambda _, j: j % 2 == 0, shape=(1, dim))
        pnt = pos / (10000 ** (idx / dim))
        self.register_buffer('psne_layer', torch.tensor(np.sin(pnt) * mask + np.cos(pnt) * (1 - mask)).float())
        self.drop_layer = nn.Dropout(drop)


def _pack_bits(mask):
    """(B, Lq, Lk) bool (True = excluded) -> (B, Lq, ceil(Lk/32)) int32 words, bit i of word w = key 32 w + i."""
    B, Lq, Lk = mask.shape
    W = (Lk + 31) // 32
    m = torch.zeros((B, Lq, W * 32), dtype=torch.int64, device=mask.device)
    m[:, :, :Lk] = mask.long()
    words = (m.view(B, Lq, W, 32) << torch.arange(32, device=mask.device)).sum(-1)
    return (words & 0xffffffff).to(torch.int64).sub_((words >> 31 & 1) << 32).to(torch.int32).contiguous()


class CaptionTransformerB200(nn.Module):
    """Constructor arguments and state_dict keys of `CaptionTransformer` (caption_tranformer.py:20-36)."""

    def __init__(self, nb_layers, input_dim, hidden_dim, ff_dim, nb_heads, drop_val, pre_norm, seq_length, nb_tokens, **kw):
        super().__init__()
        assert hidden_dim % nb_heads == 0 and (hidden_dim // nb_heads) % 4 == 0
        self.nb_heads, self.hidden_dim, self.pre_norm = nb_heads, hidden_dim, pre_norm
        self.adapter = nn.Linear(input_dim, hidden_dim) if input_dim != hidden_dim else nn.Identity()
        self.position_encoder = _Positions(seq_length, hidden_dim, drop_val)
        self.transformer_decoder = _Decoder(nb_layers, hidden_dim, ff_dim, nb_heads, drop_val, pre_norm)
        self.generator = nn.Linear(hidden_dim, nb_tokens)
        self._kern = None            # (runtime owner, tf32 flag): set by the head

    # ---- plumbing
    def bind(self, head):
        self._head = [head]          # in a list: not a submodule

    def _k(self, device, exact):
        head = self._head[0]
        tf32 = (not exact) and head.train_precision == 'tf32' and torch.is_grad_enabled()
        return _K(head._runtime(device), tf32=tf32)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None,
                exact=False):
        """tgt (B, L, hidden) token embeddings, memory (B, Q, input_dim) query embeddings; tgt_key_padding_mask (B, L) bool
        (True = padding).  Returns (list of the nb_layers block outputs (B, L, hidden), logits of the last one (B, L, vocab))
        like caption_tranformer.py:38-43.  `exact`: fp32 FMA contractions whatever the training precision (beam search)."""
        if memory_mask is not None or memory_key_padding_mask is not None or tgt_mask is not None:
            raise _lib.CggError('only the causal tgt mask + tgt_key_padding_mask form the reference uses is built')
        if not tgt.is_cuda:
            raise _lib.CggError('the caption transformer runs on CUDA only (no CPU fallback)')
        k = self._k(tgt.device, exact)
        B, L, C = tgt.shape
        Q = memory.shape[1]
        H, d = self.nb_heads, C // self.nb_heads
        scale = 1.0 / math.sqrt(d)
        active = lambda m: self.training and isinstance(m, nn.Dropout) and m.p > 0                         # noqa: E731
        drop = lambda x, m: m(x) if active(m) else x                                                       # noqa: E731
        lin = lambda x, m, res=None, relu=False: _Linear.apply(k, x, m.weight, m.bias, res, 1.0, relu)     # noqa: E731
        # residual + projection: fused into the GEMM epilogue unless a dropout sits between them (training with drop_val > 0)
        proj = lambda o, m, res, dl: (res + drop(lin(o, m), dl)) if active(dl) else lin(o, m, res=res)      # noqa: E731
        ln = lambda x, m: x if isinstance(m, nn.Identity) else _LayerNorm.apply(k, x, m.weight, m.bias, m.eps)   # noqa: E731
        mem = memory.float().contiguous().view(B * Q, -1)
        if not isinstance(self.adapter, nn.Identity):
            mem = lin(mem, self.adapter)
        x = _AddRows.apply(k, tgt.float().contiguous(), self.position_encoder.psne_layer[:L].contiguous(), B)
        x = drop(x, self.position_encoder.drop_layer).view(B * L, C)
        # causal mask (build_mask: key j > query i excluded) + key padding, as the attention kernels' bitmap
        causal = torch.ones((L, L), dtype=torch.bool, device=tgt.device).triu(1)[None].expand(B, L, L)
        if tgt_key_padding_mask is not None:
            causal = causal | tgt_key_padding_mask.bool()[:, None, :]
        self_bits = _pack_bits(causal)
        outs = []
        for blk in self.transformer_decoder.decoders:
            nz = blk.layer_normalz
            # masked self-attention (transformers.py:102-134): fused qkv, per head [q | k | v] of 3*d columns
            tmp = ln(x, nz['mha'][0])
            qkv = lin(tmp, blk.mha_layer.qkv_layer).view(B, L, H, 3, d)
            o = _AttentionViews.apply(k, qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2], self_bits, scale)
            x = ln(proj(o.view(B * L, C), blk.mha_layer.out_layer, tmp, blk.dropout_layer['mha']), nz['mha'][1])
            # cross-attention over the (adapted) query embeddings (:58-100)
            tmp = ln(x, nz['crx'][0])
            ca = blk.crx_layer
            q4 = lin(tmp, ca.to_qry).view(B, L, H, d)
            k4 = lin(mem, ca.to_key).view(B, Q, H, d)
            v4 = lin(mem, ca.to_val).view(B, Q, H, d)
            o = _AttentionViews.apply(k, q4, k4, v4, None, scale)
            x = ln(proj(o.view(B * L, C), ca.to_out, tmp, blk.dropout_layer['crx']), nz['crx'][1])
            # FFN (:27-56; note the reference feeds `agg`, not the pre-normed tmp, :229)
            tmp = ln(x, nz['ffn'][0])
            f0, f1 = blk.ffn_layer.linears[0], blk.ffn_layer.linears[1]
            if self.training and isinstance(f0[1], nn.Dropout):
                h = F.relu(f0[1](lin(x, f0[0])))                       # Linear -> Dropout -> ReLU
            else:
                h = lin(x, f0[0], relu=True)
            x = ln(proj(h, f1[0], tmp, blk.dropout_layer['ffn']), nz['ffn'][1])
            outs.append(x.view(B, L, C))
        logits = lin(x, self.generator).view(B, L, -1)
        return outs, logits


def caption_generation_loss(head, cls_emb_preds, gt_caption_ids_list, gt_caption_embs_list, gt_caption_mask_list,
                            gt_caption_nouns_ids_list=None, loss_weight=2.0):
    """loss_caption_generation of one head call (mask2former_head.py:552-583): teacher-forced next-token cross entropy of
    the caption transformer over the query embeddings, ignore_index 0, mean over all B*(T-1) positions."""
    embs = torch.stack(list(gt_caption_embs_list), 0)
    masks = torch.stack(list(gt_caption_mask_list), 0).bool()
    logits = head.caption_generator(tgt=embs[:, :-1, :], memory=cls_emb_preds,
                                    tgt_key_padding_mask=torch.logical_not(masks[:, :-1]))[1].flatten(0, 1)
    ids_list = [t.clone() for t in gt_caption_ids_list]
    if head.gen_only_obj_nouns or head.gen_mask_obj_nouns or head.gen_replace_obj_nouns:      # :563-579, host-side like the reference
        for i, ids in enumerate(ids_list):
            nouns = gt_caption_nouns_ids_list[i].cpu().numpy().tolist()
            for j in range(len(ids)):
                if int(ids[j]) not in nouns:
                    if head.gen_only_obj_nouns:
                        ids[j] = 0
                else:
                    if head.gen_mask_obj_nouns:
                        ids[j] = 0
                        break
                    if head.gen_replace_obj_nouns:
                        ids[j] = 4874
    gt = torch.stack(ids_list, 0)[:, 1:].flatten(0, 1).long()
    from .matching import _WeightedCE
    cw = torch.ones((logits.shape[1],), dtype=torch.float32, device=logits.device)
    cw[0] = 0.0                                                       # ignore_index = 0: those rows contribute nothing
    row_loss, _ = _WeightedCE.apply(logits, gt, cw)
    return loss_weight * row_loss.sum() / gt.numel()


@torch.no_grad()
def beam_search(head, memory, BOS=BOS_TOKEN, EOS=EOS_TOKEN, max_len=35, beam_width=7, alpha=0.7, tokenizer=None):
    """inference.py:84-157 with the same candidate bookkeeping: returns (best sentence token ids, score, all finished
    (ids, score) pairs); with a `tokenizer` (bert-base-uncased) also the decoded sentence as the reference returns it."""
    gen = head.caption_generator
    dev = memory.device

    def embed(ids):
        be = head.bert_embeddings
        return head.extract_word_embeddings(be.word_embeddings.weight, be.LayerNorm.weight, be.LayerNorm.bias, ids.to(dev),
                                            eps=be.LayerNorm.eps)

    def step_logits(batch_emb, mem):
        outs = gen(batch_emb, mem, exact=True)[0]
        k = gen._k(dev, True)
        rows = torch.stack([o[:, -1, :] for o in outs], 0)                                   # (layers, n, C)
        nl, n, C = rows.shape
        lg = _Linear.apply(k, rows.reshape(nl * n, C).contiguous(), gen.generator.weight, gen.generator.bias, None, 1.0, False)
        return lg.view(nl, n, -1).mean(0)                                                    # mean over the blocks' logits

    target = torch.tensor([[BOS]])
    logits = step_logits(embed(target).view(1, 1, -1), memory)[0]
    scaled = torch.log_softmax(logits[None, :], dim=1).cpu().squeeze(0)
    weights, candidates = torch.topk(scaled, k=beam_width, largest=True)
    finished, active = [], [torch.cat([target, torch.tensor([[int(i)]])], dim=1) for i in candidates]
    max_idx = 0
    while True:
        max_score, max_idx = -100, 0      # reset on EVERY pass, as the reference does (inference.py:117-118): the sentence
        #                                   returned is the best one finished in the last pass (the first one if none was)
        batch = torch.vstack(active)
        n = batch.shape[0]
        mem = torch.cat([m.repeat(n, 1, 1) for m in memory], dim=0)
        scaled = torch.log_softmax(step_logits(embed(batch).view(n, batch.shape[1], -1), mem), dim=1).cpu()
        length, vocab = batch.shape[1], scaled.shape[1]
        weighted = (scaled + weights[:, None]) / length ** alpha
        weights, candidates = torch.topk(torch.flatten(weighted), k=beam_width, largest=True)
        weights = weights * length ** alpha
        w_next, s_next, stop = [], [], False
        for idx, pos in enumerate(candidates):
            row = int(torch.div(pos, vocab, rounding_mode='floor'))
            col = int(pos % vocab)
            seq = torch.cat([active[row], torch.tensor([[col]])], dim=1)
            if col == EOS:
                flat = torch.flatten(seq).tolist()
                score = weights[idx] / len(flat) ** alpha
                finished.append((flat, float(score)))
                if score > max_score:
                    max_score, max_idx = score, len(finished) - 1
                if len(finished) == beam_width:
                    stop = True
                    break
            elif seq.shape[1] < max_len - 1:
                w_next.append(weights[row])
                s_next.append(seq)
        if stop or not s_next:
            break
        weights, active = torch.tensor(w_next), s_next
    best = finished[max_idx] if finished else (None, None)
    text = None
    if tokenizer is not None and finished:
        text = tokenizer.decode(best[0])[1:-1]                                               # inference.py:149-154
    return dict(ids=best[0], score=best[1], finished=finished, text=text)
