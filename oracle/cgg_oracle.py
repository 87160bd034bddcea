"""TEST INFRASTRUCTURE ONLY (the oracle) -- never imported by the product path.

CPU fp32 restatement, in plain torch tensor arithmetic, of the reference's decoder-head
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file, and only as the
checker / the CPU baseline.

Parity pinning: the reference ships no tests, golden vectors or checkpoints for this
path, so this restatement is pinned against OUTPUTS OF THE
REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports the unmodified reference head
(``oracle/ref_shim.py``) in the build container, runs it on seeded inputs and commits the
results under ``tests/golden/``; ``tests/test_oracle.py`` checks this file against those
vectors (and, when /root/reference is present, against the live reference).

Every function cites the reference lines it follows (paths relative to the reference
root; "mmcv"/"mmdet" = un-vendored mmcv-full 1.7.1 / mmdet 2.28.2 semantics restated in
SURVEY.md Appendix B).

All tensors are batch-first here: x (B,Q,C), memories (B,C,h,w), masks (B,Q,K) bool with
True = "do not attend" exactly as in the reference.
"""
import math

import torch

NUM_HEADS = 8


# ----------------------------------------------------------------------------- pieces
def sine_pos_enc(h, w, num_feats=128, temperature=10000.0, eps=1e-6):
    """mmdet SinePositionalEncoding(num_feats, normalize=True) of an all-valid (h,w) map
    -> (h*w, 2*num_feats), row-major keys.  Built at open_set/models/mask2former_head.py:132,
    used :798-804.  y=(row+1)/(h+eps)*2pi, x likewise; channel c of each half uses
    dim_t[c]=T^(2*(c//2)/num_feats); even c -> sin, odd c -> cos; cat(pos_y, pos_x)."""
    scale = 2 * math.pi
    rows = torch.arange(1, h + 1, dtype=torch.float32)
    cols = torch.arange(1, w + 1, dtype=torch.float32)
    y = rows / (torch.tensor(float(h)) + eps) * scale
    x = cols / (torch.tensor(float(w)) + eps) * scale
    i = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode='floor') / num_feats)
    py = y[:, None] / dim_t  # (h, F)
    px = x[:, None] / dim_t  # (w, F)
    even = (torch.arange(num_feats) % 2 == 0)
    py = torch.where(even, py.sin(), py.cos())
    px = torch.where(even, px.sin(), px.cos())
    pos = torch.cat([py[:, None, :].expand(h, w, num_feats), px[None, :, :].expand(h, w, num_feats)], dim=2)
    return pos.reshape(h * w, 2 * num_feats).contiguous()


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def linear(x, w, b):
    return x @ w.t() + b


def mha(q_in, k_in, v_in, in_w, in_b, out_w, out_b, masked=None, nheads=NUM_HEADS):
    """torch.nn.MultiheadAttention(256, 8) forward as the mmcv wrapper calls it
    (mask2former_head.py:829-840 -> mmcv MultiheadAttention -> torch MHA):
    separate q/k/v in-projections (rows 0:C / C:2C / 2C:3C of in_proj_weight),
    q scaled by 1/sqrt(d), scores masked to -inf where ``masked`` is True, softmax over
    keys, PV, out-projection.  q_in (B,Q,C); k_in, v_in (B,K,C); masked (B,Q,K) bool,
    shared by all heads (the reference replicates one mask per head, :756-757)."""
    B, Q, C = q_in.shape
    K = k_in.shape[1]
    d = C // nheads
    q = linear(q_in, in_w[:C], in_b[:C]) * (1.0 / math.sqrt(d))
    k = linear(k_in, in_w[C:2 * C], in_b[C:2 * C])
    v = linear(v_in, in_w[2 * C:], in_b[2 * C:])
    q = q.view(B, Q, nheads, d).transpose(1, 2)
    k = k.view(B, K, nheads, d).transpose(1, 2)
    v = v.view(B, K, nheads, d).transpose(1, 2)
    s = q @ k.transpose(-1, -2)  # (B,h,Q,K)
    if masked is not None:
        s = s.masked_fill(masked[:, None], float('-inf'))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Q, C)
    return linear(o, out_w, out_b)


def bilinear_resize(x, out_hw):
    """F.interpolate(mode='bilinear', align_corners=False) (mask2former_head.py:749-753)
    written out: src=(dst+0.5)*in/out-0.5 clamped at 0, neighbour index clamped at in-1,
    value = h0*(w0*a + w1*b) + h1*(w0*c + w1*d)  (the order of the CUDA kernel)."""
    B, Q, H, W = x.shape
    oh, ow = out_hw

    def axis(inn, out):
        scale = torch.tensor(inn / out, dtype=torch.float32)
        src = (torch.arange(out, dtype=torch.float32) + 0.5) * scale - 0.5
        src = torch.clamp(src, min=0.0)
        i0 = src.floor().to(torch.long).clamp(max=inn - 1)
        i1 = torch.clamp(i0 + 1, max=inn - 1)
        l1 = src - i0.to(torch.float32)
        dev = x.device
        return i0.to(dev), i1.to(dev), (1.0 - l1).to(dev), l1.to(dev)

    r0, r1, h0, h1 = axis(H, oh)
    c0, c1, w0, w1 = axis(W, ow)
    a = x[:, :, r0][:, :, :, c0]
    b = x[:, :, r0][:, :, :, c1]
    c = x[:, :, r1][:, :, :, c0]
    d = x[:, :, r1][:, :, :, c1]
    top = w0 * a + w1 * b
    bot = w0 * c + w1 * d
    return h0[:, None] * top + h1[:, None] * bot


def attn_mask_from_logits(mask_pred, target_hw):
    """mask2former_head.py:749-759: bilinear resize to the level size, then
    ``sigmoid() < 0.5`` (True = masked).  One mask per (image, query); identical across
    heads.  Returns (B,Q,K) bool."""
    d = bilinear_resize(mask_pred, target_hw)
    return (torch.sigmoid(d) < 0.5).flatten(2)


def apply_fallback(masked):
    """mask2former_head.py:825-826: a row that masks every key is cleared."""
    all_masked = masked.sum(-1) == masked.shape[-1]
    return masked & ~all_masked[..., None]


def head_call(sd, x, mask_features, target_hw, pred_emb_norm=False):
    """forward_head, mask2former_head.py:711-761.  x (B,Q,C) (un-normalised decoder state).
    Returns cls (B,Q,ncls+1), emb (B,Q,768), mask_pred (B,Q,H4,W4), masked (B,Q,K) bool."""
    z = layer_norm(x, sd['transformer_decoder.post_norm.weight'], sd['transformer_decoder.post_norm.bias'])
    cls = linear(z, sd['cls_embed.weight'], sd['cls_embed.bias'])
    emb = linear(z, sd['v2l_transform.weight'], sd['v2l_transform.bias'])
    if pred_emb_norm:
        emb = emb / emb.norm(dim=-1, keepdim=True)
    m = torch.relu(linear(z, sd['mask_embed.0.weight'], sd['mask_embed.0.bias']))
    m = torch.relu(linear(m, sd['mask_embed.2.weight'], sd['mask_embed.2.bias']))
    m = linear(m, sd['mask_embed.4.weight'], sd['mask_embed.4.bias'])
    B, C, H, W = mask_features.shape
    mask_pred = (m @ mask_features.reshape(B, C, H * W)).reshape(B, -1, H, W)
    masked = attn_mask_from_logits(mask_pred, target_hw)
    return cls, emb, mask_pred, masked, m


def decoder_layer(sd, i, x, qe, key_in, val_in, masked):
    """One DetrTransformerDecoderLayer, operation_order (cross_attn, norm, self_attn, norm,
    ffn, norm) -- mmcv BaseTransformerLayer semantics, called at mask2former_head.py:829-840.
    x (B,Q,C); qe (Q,C) query_embed; key_in = mem+level_embed+pos (B,K,C); val_in =
    mem+level_embed (B,K,C); masked (B,Q,K) bool after the fallback."""
    p = 'transformer_decoder.layers.%d.' % i
    a0, a1 = p + 'attentions.0.attn.', p + 'attentions.1.attn.'
    x = x + mha(x + qe, key_in, val_in, sd[a0 + 'in_proj_weight'], sd[a0 + 'in_proj_bias'],
                sd[a0 + 'out_proj.weight'], sd[a0 + 'out_proj.bias'], masked)
    x = layer_norm(x, sd[p + 'norms.0.weight'], sd[p + 'norms.0.bias'])
    x = x + mha(x + qe, x + qe, x, sd[a1 + 'in_proj_weight'], sd[a1 + 'in_proj_bias'],
                sd[a1 + 'out_proj.weight'], sd[a1 + 'out_proj.bias'], None)
    x = layer_norm(x, sd[p + 'norms.1.weight'], sd[p + 'norms.1.bias'])
    f = torch.relu(linear(x, sd[p + 'ffns.0.layers.0.0.weight'], sd[p + 'ffns.0.layers.0.0.bias']))
    x = x + linear(f, sd[p + 'ffns.0.layers.1.weight'], sd[p + 'ffns.0.layers.1.bias'])
    return layer_norm(x, sd[p + 'norms.2.weight'], sd[p + 'norms.2.bias'])


def decoder_forward(sd, mask_features, memories, num_layers=9, pred_emb_norm=False,
                    teacher_x=None, forced_masked=None):
    """Mask2FormerHeadOpen.forward after the pixel decoder, mask2former_head.py:787-849.

    sd: state_dict (reference key names).  memories: [mem32, mem16, mem8] each (B,C,h,w).
    teacher_x: optional list of (B,Q,C) states; when given, layer i consumes teacher_x[i]
    instead of its own running state (teacher-forced comparison).
    forced_masked: optional list of boolean attention masks (before the fallback) used INSTEAD of the ones derived from
    this run's own logits (they are detached constants of the graph, :759) -- lets a reduced-precision run of this same
    function be compared with the fp32 one on identical masks.

    Returns dict with lists of length num_layers+1: cls, emb, mask; plus x (decoder state fed
    to each head call), masked (bool mask produced by each head call, BEFORE fallback),
    mask_embed."""
    B = mask_features.shape[0]
    C = mask_features.shape[1]
    L = len(memories)
    key_in, val_in, sizes = [], [], []
    for l, mem in enumerate(memories):
        h, w = mem.shape[-2:]
        flat = mem.flatten(2).transpose(1, 2) + sd['level_embed.weight'][l]  # :792-796
        pos = sine_pos_enc(h, w, C // 2).to(mem.device)                      # :798-804
        val_in.append(flat)
        key_in.append(flat + pos)
        sizes.append((h, w))
    qe = sd['query_embed.weight']
    x = sd['query_feat.weight'][None].expand(B, -1, -1).contiguous()        # :808-811
    out = dict(cls=[], emb=[], mask=[], x=[], masked=[], mask_embed=[])

    def call_head(x, lvl):
        cls, emb, mp, masked, me = head_call(sd, x, mask_features, sizes[lvl], pred_emb_norm)
        out['cls'].append(cls), out['emb'].append(emb), out['mask'].append(mp)
        out['x'].append(x), out['masked'].append(masked), out['mask_embed'].append(me)
        if forced_masked is not None:
            return forced_masked[len(out['masked']) - 1]
        return masked

    masked = call_head(x, 0)                                                # :816-820
    for i in range(num_layers):                                             # :822-847
        lvl = i % L
        if teacher_x is not None:
            x = teacher_x[i]
            masked = head_call(sd, x, mask_features, sizes[lvl], pred_emb_norm)[3]
        x = decoder_layer(sd, i, x, qe, key_in[lvl], val_in[lvl], apply_fallback(masked))
        masked = call_head(x, (i + 1) % L)
    return out


# ----------------------------------------------------------------------- grounding side
def cls_emb_logits(emb, class_embs, temperature=10.0):
    """_get_cls_emb_logits, mask2former_head.py:631-648."""
    return emb @ class_embs.t() / temperature


def noun_embeddings(table, ln_w, ln_b, ids, text_emb_norm=True, eps=1e-12):
    """extract_word_embeddings (bert branch), mask2former_head.py:686-698 with
    BertEmbeddings (models/utils/bert_embeddings.py:4-13): table lookup + BERT LayerNorm."""
    e = table[ids]
    return layer_norm(e, ln_w, ln_b, eps) if text_emb_norm else e


def test_time_grounding(emb_last, noun_embs):
    """simple_test ``att``, mask2former_head.py:973-978: query x noun similarity."""
    return emb_last @ noun_embs.t()


def grounding_loss(pred, cap, cap_mask, temperature=10.0, loss_weight=1.0):
    """losses/grounding_loss.py:9-77 restated as ONE similarity contraction
    S[i,j,t,q] = cap[i,t].pred[j,q]/T over all (caption i, image j) pairs instead of the
    reference's B^2-fold operand replication (:23-30).  Quirks kept: the token mask weights
    only the l2v attention (:42); the v2l softmax runs over all max_tokens incl. padding
    (:40,:47); divisor max(num_tokens,1) (:45); empty captions -> max().detach()+100
    (:52-61); cost rows = captions, cols = images (:63,:69).
    pred (B,Q,D); cap (B,Tk,D); cap_mask (B,Tk) {0,1}."""
    B, Q, D = pred.shape
    maskf = cap_mask.to(pred.dtype)
    ntok = cap_mask.sum(1)
    S = torch.einsum('itd,jqd->ijtq', cap, pred) / temperature
    dist = -S
    a_l2v = torch.softmax(S, dim=3) * maskf[:, None, :, None]
    a_v2l = torch.softmax(S, dim=2)
    denom = torch.clamp(ntok, min=1).to(pred.dtype)
    g_l2v = (a_l2v * dist).sum(3).sum(2) / denom[:, None]
    g_v2l = (a_v2l * dist).sum(3).sum(2) / Q
    has = (ntok > 0)[:, None]
    g_l2v = torch.where(has, g_l2v, g_l2v.max().detach() + 100.0)
    g_v2l = torch.where(has, g_v2l, g_v2l.max().detach() + 100.0)

    def sym_ce(cost):
        l_cap = -torch.log_softmax(-cost, dim=0).diagonal().mean()
        l_img = -torch.log_softmax(-cost, dim=1).diagonal().mean()
        return l_cap + l_img

    return loss_weight * (sym_ce(g_l2v) + sym_ce(g_v2l)) / 4


def pack_mask_bits(masked):
    """(B,Q,K) bool -> (B,Q,ceil(K/32)) int32 words, bit k%32 of word k//32 = masked[...,k]
    (the bitmap layout of include/cgg_b200.h); tail bits of the last word are 0."""
    B, Q, K = masked.shape
    W = (K + 31) // 32
    pad = torch.zeros(B, Q, W * 32, dtype=torch.int64)
    pad[..., :K] = masked.to(torch.int64)
    weights = (1 << torch.arange(32, dtype=torch.int64))
    words = (pad.view(B, Q, W, 32) * weights).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    return words.to(torch.int32)
