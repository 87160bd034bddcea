"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/open_set/models/mask2former_head.py and losses/grounding_loss.py, loaded
through oracle/ref_shim.py) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures hold seeds + reference outputs, never weights: weights and inputs are
re-drawn from the same CPU generators (cgg_b200/synth.py) by the tests, and a parameter
checksum in each fixture detects RNG drift.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, cgg_oracle as O  # noqa: E402
from cgg_b200 import synth  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cases import (HEAD_CASES, GROUNDING_CASES, MASK_SAMPLE_STRIDE, GRAD_SAMPLE_STRIDE, case_tensors,  # noqa: E402
                   param_checksum, grounding_tensors)

HERE = os.path.dirname(os.path.abspath(__file__))

def run_head_case(name, c):
    R = ref_shim.REF_ROOT
    if c.get('class_embs') == 'coco_panoptic_p20':
        head = ref_shim.build_reference_head(num_queries=c['num_queries'], num_known=64, num_stuff=53,
                                             unknown_file=R + '/datasets/unknown/unknown_p20.txt',
                                             class_to_emb_file=R + '/datasets/embeddings/coco_panoptic_class_with_bert_emb.json')
    else:
        head = ref_shim.build_reference_head(num_queries=c['num_queries'],
                                             known_file=R + '/datasets/unknown/known_65.txt',
                                             unknown_file=R + '/datasets/unknown/unknown_17.txt')
    sd, mf, mems = case_tensors(c)
    if c.get('class_embs'):      # the committed fixture IS what the reference loader builds from its json files
        assert torch.equal(head.class_embs, sd['class_embs'])
    head.load_state_dict(sd, strict=True)
    masks = []
    orig = head.forward_head

    def rec(decoder_out, mask_feature, size):
        r = orig(decoder_out, mask_feature, size)
        B = mask_feature.shape[0]
        # (B*heads, Q, K) -> one copy per image (all heads identical, :756-757)
        am = r[3].view(B, head.num_heads, r[3].shape[1], -1)
        assert bool((am == am[:, :1]).all())
        masks.append(am[:, 0].clone())
        return r

    head.forward_head = rec
    cls, emb, mask = ref_shim.run_reference_head(head, mf, mems)
    out = dict(param_checksum=np.float64(param_checksum(sd)))
    n_fallback = 0
    for j in range(len(cls)):
        out['cls_%d' % j] = cls[j].numpy()
        out['emb_%d' % j] = emb[j].numpy() if j in (0, 4, 9) else emb[j].numpy()[:, :, ::16]
        flat = mask[j].flatten()
        out['mask_sample_%d' % j] = flat[::MASK_SAMPLE_STRIDE].numpy()
        out['mask_sum_%d' % j] = np.float64(mask[j].double().sum())
        out['mask_abssum_%d' % j] = np.float64(mask[j].double().abs().sum())
        out['bits_%d' % j] = O.pack_mask_bits(masks[j]).numpy()
        n_fallback += int((masks[j].sum(-1) == masks[j].shape[-1]).sum())
    out['n_all_masked_rows'] = np.int64(n_fallback)
    out['last_mask_full'] = mask[-1].numpy()
    if c.get('class_embs'):
        with torch.no_grad():
            out['cls_emb_logits_9'] = head._get_cls_emb_logits(emb[9]).numpy()
    np.savez_compressed(os.path.join(HERE, 'head_%s.npz' % name), **out)
    dens = [float(m.float().mean()) for m in masks]
    print(name, 'all-masked rows:', n_fallback, 'mask density per call:', ['%.2f' % d for d in dens])


def run_grounding():
    _, gl = ref_shim.load_reference_modules()
    out = {}
    for name in GROUNDING_CASES:
        pred, cap, m = grounding_tensors(name)
        pred.requires_grad_(True)
        loss = gl.GroundingLoss(loss_weight=2.0)(pred, cap, m, 10.0)
        loss.backward()
        out[name + '_loss'] = np.float32(loss.item())
        out[name + '_grad_sample'] = pred.grad.flatten()[::GRAD_SAMPLE_STRIDE].numpy()
        out[name + '_grad_abssum'] = np.float64(pred.grad.double().abs().sum())
        print('grounding', name, loss.item())
    np.savez_compressed(os.path.join(HERE, 'grounding.npz'), **out)


def run_embeddings():
    """extract_word_embeddings + _get_cls_emb_logits + test-time att, on the real head
    object with a seeded 400-row BERT-like table."""
    head_mod, _ = ref_shim.load_reference_modules()
    R = ref_shim.REF_ROOT
    head = ref_shim.build_reference_head(num_queries=16,
                                         known_file=R + '/datasets/unknown/known_65.txt',
                                         unknown_file=R + '/datasets/unknown/unknown_17.txt')
    sd = synth.make_params(seed=9, num_queries=16)
    head.load_state_dict(sd, strict=True)
    ids, mask, table, lw, lb = synth.make_captions(5, 4, vocab=400 + 1000)
    g = torch.Generator().manual_seed(11)
    lw = lw + 0.1 * torch.randn(lw.shape, generator=g)
    lb = lb + 0.1 * torch.randn(lb.shape, generator=g)
    emb_layer = torch.nn.Embedding(table.shape[0], 768, padding_idx=0)
    emb_layer.weight.data.copy_(table)
    cfg = types.SimpleNamespace(vocab_size=table.shape[0], hidden_size=768, pad_token_id=0, layer_norm_eps=1e-12)
    ln = torch.nn.LayerNorm(768, eps=1e-12)
    ln.weight.data.copy_(lw), ln.bias.data.copy_(lb)
    fake_bert = types.SimpleNamespace(config=cfg, embeddings=types.SimpleNamespace(word_embeddings=emb_layer, LayerNorm=ln))
    head.bert_embeddings = head_mod.BertEmbeddings(fake_bert)
    with torch.no_grad():
        embs, _ = head.extract_word_embeddings([i for i in ids], [m for m in mask])
        pred = torch.randn((4, 16, 768), generator=g)
        logits = head._get_cls_emb_logits(pred)
        att = torch.matmul(pred[0], embs[0].t())
    np.savez_compressed(os.path.join(HERE, 'embeddings.npz'), noun_embs=torch.stack(embs).numpy(),
                        ln_w=lw.numpy(), ln_b=lb.numpy(), pred=pred.numpy(), logits=logits.numpy(),
                        att=att.numpy(), param_checksum=np.float64(param_checksum(sd)))
    print('embeddings ok')


if __name__ == '__main__':
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for n, c in HEAD_CASES.items():
        if not only or n in only:
            run_head_case(n, c)
    if not only:
        run_grounding()
        run_embeddings()
