"""CPU study: which bf16 roundings of the throughput mode drive the free-running end-to-end error
(run in the build container; uses the oracle, never imported by the product).

Each flag turns ONE class of roundings on in an otherwise fp32 restatement of the path:
  w    small-M weights (q/out/self-qkv/ffn) rounded to bf16
  act  small-M GEMM activations (x+qe, x1+qe, x2, attention outputs) rounded to bf16
  ffn  FFN hidden activations rounded to bf16
  kvw  K/V projection weights rounded to bf16
  kv   projected K and V stored as bf16
  q    attention Q rounded to bf16
  p    attention probabilities (unnormalised exp) rounded to bf16
  me   mask embeddings rounded to bf16 for the mask einsum (+ bf16 output)
Prints final-call max-abs error / logit range for mask and emb, and per-call mask-bit agreement.
"""
import itertools
import math
import sys

import torch

sys.path.insert(0, '.')
from oracle import cgg_oracle as O
from cgg_b200 import synth


def r(x, on):
    return x.bfloat16().float() if on else x


def mha(q_in, k, v, in_w, in_b, out_w, out_b, masked, f, kv_ready=False, k_in=None, v_in=None):
    B, Q, C = q_in.shape
    d = C // 8
    q = O.linear(r(q_in, 'act' in f), r(in_w[:C], 'w' in f), in_b[:C]) * (1.0 / math.sqrt(d))
    if not kv_ready:
        k = O.linear(r(k_in, 'act' in f), r(in_w[C:2 * C], 'w' in f), in_b[C:2 * C])
        v = O.linear(r(v_in, 'act' in f), r(in_w[2 * C:], 'w' in f), in_b[2 * C:])
        k, v = r(k, 'kv' in f), r(v, 'kv' in f)
    K = k.shape[1]
    q = r(q, 'q' in f).view(B, Q, 8, d).transpose(1, 2)
    k = k.view(B, K, 8, d).transpose(1, 2)
    v = v.view(B, K, 8, d).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if masked is not None:
        s = s.masked_fill(masked[:, None], float('-inf'))
    m = s.max(-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)
    o = (r(p, 'p' in f) @ v) / l
    o = o.transpose(1, 2).reshape(B, Q, C)
    return O.linear(r(o, 'act' in f), r(out_w, 'w' in f), out_b)


def forward(sd, mf, mems, f):
    B, C = mf.shape[:2]
    qe = sd['query_embed.weight']
    x = sd['query_feat.weight'][None].expand(B, -1, -1).contiguous()
    sizes = [tuple(m.shape[-2:]) for m in mems]
    ks, vs = [], []
    for i in range(9):
        l = i % 3
        mem = mems[l]
        h, w = sizes[l]
        flat = mem.flatten(2).transpose(1, 2)
        pos = O.sine_pos_enc(h, w, C // 2)
        p = 'transformer_decoder.layers.%d.attentions.0.attn.' % i
        W, b = sd[p + 'in_proj_weight'], sd[p + 'in_proj_bias']
        lev = sd['level_embed.weight'][l]
        Wk, Wv = r(W[C:2 * C], 'kvw' in f), r(W[2 * C:], 'kvw' in f)
        # K = mem Wk^T (rounded) + exact table; V = mem Wv^T + bias (rounded)
        k = r(flat @ Wk.t(), 'kv' in f) + ((pos + lev) @ W[C:2 * C].t() + b[C:2 * C])
        v = r(flat @ Wv.t() + (lev @ W[2 * C:].t() + b[2 * C:]), 'kv' in f)
        ks.append(k), vs.append(v)
    out = dict(mask=[], emb=[], masked=[], cls=[])

    def head(x, lvl):
        cls, emb, mp, masked, me = O.head_call(sd, x, mf, sizes[lvl])
        if 'me' in f:
            mp = r((r(me, True) @ mf.reshape(B, C, -1)).reshape(mp.shape), True)
        out['mask'].append(mp), out['emb'].append(emb), out['masked'].append(masked), out['cls'].append(cls)
        return masked

    masked = head(x, 0)
    for i in range(9):
        p = 'transformer_decoder.layers.%d.' % i
        a0, a1 = p + 'attentions.0.attn.', p + 'attentions.1.attn.'
        x = x + mha(x + qe, ks[i], vs[i], sd[a0 + 'in_proj_weight'], sd[a0 + 'in_proj_bias'], sd[a0 + 'out_proj.weight'],
                    sd[a0 + 'out_proj.bias'], O.apply_fallback(masked), f, kv_ready=True)
        x = O.layer_norm(x, sd[p + 'norms.0.weight'], sd[p + 'norms.0.bias'])
        x = x + mha(x + qe, None, None, sd[a1 + 'in_proj_weight'], sd[a1 + 'in_proj_bias'], sd[a1 + 'out_proj.weight'],
                    sd[a1 + 'out_proj.bias'], None, f, k_in=x + qe, v_in=x)
        x = O.layer_norm(x, sd[p + 'norms.1.weight'], sd[p + 'norms.1.bias'])
        hdn = torch.relu(O.linear(r(x, 'act' in f), r(sd[p + 'ffns.0.layers.0.0.weight'], 'w' in f), sd[p + 'ffns.0.layers.0.0.bias']))
        x = x + O.linear(r(hdn, 'ffn' in f), r(sd[p + 'ffns.0.layers.1.weight'], 'w' in f), sd[p + 'ffns.0.layers.1.bias'])
        x = O.layer_norm(x, sd[p + 'norms.2.weight'], sd[p + 'norms.2.bias'])
        masked = head(x, (i + 1) % 3)
    return out


def report(name, ref, got):
    rng = float(ref['mask'][9].abs().max())
    em = float((got['mask'][9] - ref['mask'][9]).abs().max()) / rng
    ee = float((got['emb'][9] - ref['emb'][9]).abs().max()) / float(ref['emb'][9].abs().max())
    ec = float((got['cls'][9] - ref['cls'][9]).abs().max()) / float(ref['cls'][9].abs().max())
    bits = min(float((got['masked'][j] == ref['masked'][j]).float().mean()) for j in range(9))
    print('%-28s mask %.2e  emb %.2e  cls %.2e  min bit agreement %.5f' % (name, em, ee, ec, bits))


if __name__ == '__main__':
    torch.set_num_threads(8)
    for (Q, B, H, W, ps, iseed) in [(100, 2, 256, 256, 33, 9), (32, 2, 256, 256, 7, 7)]:
        sd = synth.make_params(seed=ps, num_queries=Q, perturb=True)
        mf, mems = synth.make_inputs(iseed, B, H, W)
        mf, mems = mf.bfloat16().float(), [m.bfloat16().float() for m in mems]
        with torch.no_grad():
            ref = forward(sd, mf, mems, set())
            chk = O.decoder_forward(sd, mf, mems)
            report('restatement vs oracle', chk, ref)
            allf = ['w', 'act', 'ffn', 'kvw', 'kv', 'q', 'p', 'me']
            for fl in allf:
                report(fl, ref, forward(sd, mf, mems, {fl}))
            report('all', ref, forward(sd, mf, mems, set(allf)))
            report('all - w,act,ffn', ref, forward(sd, mf, mems, set(allf) - {'w', 'act', 'ffn'}))
            report('all - w,act,ffn,q', ref, forward(sd, mf, mems, set(allf) - {'w', 'act', 'ffn', 'q'}))
            report('kvw,kv,p,me', ref, forward(sd, mf, mems, {'kvw', 'kv', 'p', 'me'}))
            report('kv,p,me', ref, forward(sd, mf, mems, {'kv', 'p', 'me'}))
        print()
