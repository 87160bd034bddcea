"""Stage-by-stage numeric check of the bf16 / tcgen05 path against plain torch ops on the same
(bf16-rounded) operands.  Not a pytest: run on the GPU box, prints granular error statistics
so that a wrong descriptor / swizzle / epilogue mapping can be localised from one run.

    python tests/tc_debug.py [H W B Q]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cgg_b200 import synth  # noqa: E402
from cgg_b200.head import build_head_from_state_dict  # noqa: E402


def ws_view(rt, name, shape, dtype):
    off = rt.lib.cgg_workspace_offset(rt.handle, rt.batch, name.encode())
    assert off != 2 ** 64 - 1, name
    n = 1
    for s in shape:
        n *= s
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    return rt.workspace[off:off + nbytes].view(dtype).view(shape)


def report(tag, got, want, detail_dims=None):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    rng = float(want.abs().max())
    print('%-28s max-abs err %.4g  (range %.4g, rel %.3g)  mean err %.3g  nan=%d'
          % (tag, float(err.max()), rng, float(err.max()) / max(rng, 1e-30), float(err.mean()), int(torch.isnan(got).sum())))
    return float(err.max()) / max(rng, 1e-30)


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    Q = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    dev = torch.device('cuda', 0)
    sd = synth.make_params(seed=3, num_queries=Q, perturb=True)
    mf, mems = synth.make_inputs(1, B, H, W)
    mf_b, mems_b = mf.to(dev).bfloat16(), [m.to(dev).bfloat16() for m in mems]
    head = build_head_from_state_dict(sd, Q, 49, 'bf16', dev)
    rt = head._runtime(dev)
    sizes = [tuple(m.shape[-2:]) for m in mems]
    H4, W4 = mf.shape[-2:]
    rt.prepare(H4, W4, sizes, B)
    C = 256
    sdd = {k: v.to(dev) for k, v in sd.items()}
    worst = 0.0

    # ---- K4: K/V projection
    rt.kv_project(mems_b)
    torch.cuda.synchronize()
    from oracle import cgg_oracle as O
    for l in range(3):
        K = sizes[l][0] * sizes[l][1]
        kv = ws_view(rt, 'kv%d' % l, (B, K, 3 * 2 * C), torch.bfloat16)
        flat = mems_b[l].float().flatten(2).transpose(1, 2)            # (B,K,C) from bf16 inputs
        pos = O.sine_pos_enc(sizes[l][0], sizes[l][1]).to(dev)
        lvl = sdd['level_embed.weight'][l]
        for sl in range(3):
            i = sl * 3 + l
            w_in = sdd['transformer_decoder.layers.%d.attentions.0.attn.in_proj_weight' % i]
            b_in = sdd['transformer_decoder.layers.%d.attentions.0.attn.in_proj_bias' % i]
            wk_b = w_in[C:2 * C].bfloat16().float()
            wv_b = w_in[2 * C:].bfloat16().float()
            k_want = flat @ wk_b.t() + ((pos + lvl) @ w_in[C:2 * C].t() + b_in[C:2 * C])
            v_want = flat @ wv_b.t() + (lvl @ w_in[2 * C:].t() + b_in[2 * C:])
            worst = max(worst, report('kv level %d layer %d K' % (l, i), kv[:, :, sl * C:(sl + 1) * C], k_want))
            worst = max(worst, report('kv level %d layer %d V' % (l, i), kv[:, :, (3 + sl) * C:(4 + sl) * C], v_want))
            if l == 0 and sl == 0:
                e = (kv[:, :, :C].float() - k_want).abs()
                print('   per-key-block(32) max err :', [round(float(x), 3) for x in e[0].amax(1).view(-1, 32).amax(1)[:8]])
                print('   per-col-block(16) max err :', [round(float(x), 3) for x in e[0].amax(0).view(-1, 16).amax(1)[:16]])

    # ---- head call 0: me -> bits + einsum
    x0 = sdd['query_feat.weight'][None].expand(B, -1, -1).contiguous()
    for lvl_idx in range(3):
        cls, emb, mask, me, bm, am = rt.head_call(x0, mf_b, lvl_idx)
        torch.cuda.synchronize()
        K = sizes[lvl_idx][0] * sizes[lvl_idx][1]
        me_b = me.bfloat16().float()
        if lvl_idx == 0:
            want = torch.einsum('bqc,bcp->bqp', me_b, mf_b.float().flatten(2)).view(B, Q, H4, W4)
            worst = max(worst, report('einsum (single call)', mask, want))
            e = (mask.float() - want).abs()[0]
            print('   per-query max err (first 16):', [round(float(x), 3) for x in e.flatten(1).amax(1)[:16]])
            print('   per-row max err (first 8 rows):', [round(float(x), 3) for x in e.amax(0).amax(1)[:8]])
        fds2 = ws_view(rt, 'fds%d' % lvl_idx, (B, 2 * C, K), torch.bfloat16)
        fds = fds2[:, :C].float() + fds2[:, C:].float()               # hi + lo
        want_ds = torch.nn.functional.interpolate(mf_b.float(), sizes[lvl_idx], mode='bilinear', align_corners=False)
        worst = max(worst, report('downsample level %d (hi+lo)' % lvl_idx, fds, want_ds.flatten(2)))
        logits = torch.einsum('bqc,bck->bqk', me, fds)                # split precision ~ fp32 operands
        want_bits = (logits.sigmoid() < 0.5)
        got = torch.zeros_like(want_bits)
        words = bm.cpu()
        for k in range(K):
            got[:, :, k] = ((words[:, :, k // 32] >> (k % 32)) & 1).bool().to(dev)
        agree = float((got == want_bits).float().mean())
        near = (logits.abs() < 1e-4)
        hard_bad = int(((got != want_bits) & ~near).sum())
        print('bits level %d: agreement %.5f, disagreements away from threshold: %d, all_masked ok: %s'
              % (lvl_idx, agree, hard_bad, bool((am.bool() == want_bits.all(-1)).all())))
        if hard_bad:
            worst = 1.0
    print('WORST_REL_ERR %.4g' % worst)
    print('TC_DEBUG', 'PASS' if worst < 2e-2 else 'FAIL')


if __name__ == '__main__':
    main()
