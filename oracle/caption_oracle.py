"""TEST INFRASTRUCTURE -- CPU restatement of the caption generator: CaptionTransformer
(open_set/models/transformers/caption_tranformer.py:20-43; blocks open_set/models/transformers/transformers.py:58-134,
:180-234, :252-267), the caption-generation loss (open_set/models/mask2former_head.py:552-583) and the beam search
(open_set/utils/eval/inference.py:84-157) in plain torch on a state_dict with the reference's key names.  Only tests/ may
import this file.  Pinned by tests/test_caption_cpu.py against the UNMODIFIED reference modules imported through
oracle/ref_shim.py (forward, loss, and the beam search with the tokenizer download stubbed out) and by the committed
fixture tests/golden/caption.npz."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def positions(seq_length, dim):
    """PositionalEncoding buffer, transformers.py:9-20."""
    pos = np.arange(0, seq_length)[:, None]
    idx = np.fromfunction(lambda _, j: j - j % 2, shape=(1, dim))
    mask = np.fromfunction(lambda _, j: j % 2 == 0, shape=(1, dim))
    pnt = pos / (10000 ** (idx / dim))
    return torch.tensor(np.sin(pnt) * mask + np.cos(pnt) * (1 - mask)).float()


def _mh(x, heads):
    b, l, c = x.shape
    return x.reshape(b, l, heads, c // heads).permute(0, 2, 1, 3)


def _attend(q, k, v, mask=None, key_padding_mask=None):
    w = q @ k.transpose(-2, -1) / np.sqrt(q.shape[-1])
    if mask is not None:
        w = w.masked_fill(mask, float('-inf'))
    if key_padding_mask is not None:
        w = w.masked_fill(key_padding_mask[:, None, None, :], float('-inf'))
    r = torch.softmax(w, dim=-1) @ v
    return torch.flatten(r.permute(0, 2, 1, 3), start_dim=2)


def forward(sd, cfg, tgt, memory, tgt_key_padding_mask=None, prefix='caption_generator.'):
    """(outputs of every block, logits of the last) -- post-norm blocks (pre_norm=False), dropout inactive (eval)."""
    g = lambda k: sd[prefix + k]                                       # noqa: E731
    lin = lambda x, k: F.linear(x, g(k + '.weight'), g(k + '.bias'))   # noqa: E731
    ln = lambda x, k: F.layer_norm(x, (x.shape[-1],), g(k + '.weight'), g(k + '.bias'), 1e-5)   # noqa: E731
    H = cfg['nb_heads']
    if cfg['input_dim'] != cfg['hidden_dim']:
        memory = lin(memory, 'adapter')
    L = tgt.shape[1]
    x = tgt + g('position_encoder.psne_layer')[:L][None]
    causal = torch.as_tensor(np.fromfunction(lambda i, j: j > i, shape=(L, L))).to(tgt.device)
    outs = []
    for i in range(cfg['nb_layers']):
        p = 'transformer_decoder.decoders.%d.' % i
        b, l, c = x.shape
        qkv = lin(x, p + 'mha_layer.qkv_layer').reshape(b, l, H, 3 * (c // H)).permute(0, 2, 1, 3)
        q, k, v = torch.chunk(qkv, 3, dim=-1)
        x = ln(x + lin(_attend(q, k, v, causal, tgt_key_padding_mask), p + 'mha_layer.out_layer'), p + 'layer_normalz.mha.1')
        q = _mh(lin(x, p + 'crx_layer.to_qry'), H)
        k = _mh(lin(memory, p + 'crx_layer.to_key'), H)
        v = _mh(lin(memory, p + 'crx_layer.to_val'), H)
        x = ln(x + lin(_attend(q, k, v), p + 'crx_layer.to_out'), p + 'layer_normalz.crx.1')
        h = F.relu(lin(x, p + 'ffn_layer.linears.0.0'))
        x = ln(x + lin(h, p + 'ffn_layer.linears.1.0'), p + 'layer_normalz.ffn.1')
        outs.append(x)
    return outs, lin(x, 'generator')


def caption_loss(sd, cfg, cls_emb_preds, caption_ids, caption_embs, caption_mask, loss_weight=2.0):
    """mask2former_head.py:552-583 (no gen_* id rewriting): CE(ignore_index=0), mean over all B*(T-1) positions."""
    logits = forward(sd, cfg, caption_embs[:, :-1, :], cls_emb_preds,
                     tgt_key_padding_mask=torch.logical_not(caption_mask.bool()[:, :-1]))[1].flatten(0, 1)
    gt = caption_ids[:, 1:].flatten(0, 1)
    return loss_weight * F.cross_entropy(logits, gt, reduction='none', ignore_index=0).mean()


def embed_ids(sd, ids):
    """get_ids_embedding, inference.py:78-83 (without its squeeze: shapes stay (n, L, d))."""
    e = F.embedding(ids, sd['bert_embeddings.word_embeddings.weight'])
    return F.layer_norm(e, (e.shape[-1],), sd['bert_embeddings.LayerNorm.weight'], sd['bert_embeddings.LayerNorm.bias'], 1e-12)


@torch.no_grad()
def beam_search(sd, cfg, memory, BOS, EOS, max_len=35, beam_width=7, alpha=0.7):
    """inference.py:84-157: returns (best sentence ids, all finished (ids, score))."""
    def step(ids, mem):
        outs = forward(sd, cfg, embed_ids(sd, ids), mem)[0]
        return torch.mean(torch.stack([F.linear(o[:, -1, :], sd['caption_generator.generator.weight'],
                                                sd['caption_generator.generator.bias']) for o in outs]), dim=0)
    target = torch.tensor([[BOS]])
    scaled = torch.log_softmax(step(target, memory), dim=1).squeeze(0)
    weights, cand = torch.topk(scaled, k=beam_width, largest=True)
    finished, active = [], [torch.cat([target, torch.tensor([[int(i)]])], dim=1) for i in cand]
    max_idx = 0
    while True:
        max_score, max_idx = -100, 0      # reset on EVERY pass, as the reference does (inference.py:117-118): the sentence
        #                                   returned is the best one finished in the last pass (the first one if none was)
        batch = torch.vstack(active)
        mem = torch.cat([m.repeat(batch.shape[0], 1, 1) for m in memory], dim=0)
        scaled = torch.log_softmax(step(batch, mem), dim=1)
        length, vocab = batch.shape[1], scaled.shape[1]
        weighted = (scaled + weights[:, None]) / length ** alpha
        weights, cand = torch.topk(torch.flatten(weighted), k=beam_width, largest=True)
        weights = weights * length ** alpha
        w_next, s_next, stop = [], [], False
        for idx, pos in enumerate(cand):
            row, col = int(torch.div(pos, vocab, rounding_mode='floor')), int(pos % vocab)
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
    return (finished[max_idx][0] if finished else None), finished
