"""CPU tests of the oracle for the caption generator (oracle/caption_oracle.py): pinned on the
UNMODIFIED reference modules -- CaptionTransformer.forward, the caption-generation cross entropy, and beam_search with its
tokenizer download stubbed out -- and on the committed fixture tests/golden/caption.npz; plus the host-side pieces of the
product (state_dict keys, the mask bit packing)."""
import os
import types
import numpy as np
import pytest
import torch

from oracle import caption_oracle as CO
from oracle import ref_shim
from cgg_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
need_ref = pytest.mark.skipif(not ref_shim.reference_available(), reason='reference tree not present')
CFG = synth.CAPTION_CFG_SMALL
BOS, EOS = 101, 102


def case(seed, B=2, Q=12, T=35):
    g = torch.Generator().manual_seed(seed)
    memory = torch.randn((B, Q, 768), generator=g)
    ids = torch.randint(200, CFG['nb_tokens'], (B, T), generator=g)
    mask = torch.zeros((B, T), dtype=torch.long)
    for b, n in enumerate([9, 20][:B]):
        mask[b, :n] = 1
        ids[b, 0], ids[b, n - 1] = BOS, EOS
    ids = ids * mask
    return memory, ids, mask


def _live(sd):
    ct, inf = ref_shim.load_caption_modules()
    m = ct.CaptionTransformer(**CFG).eval()
    own = {k[len('caption_generator.'):]: v for k, v in sd.items() if k.startswith('caption_generator.')}
    own['position_encoder.psne_layer'] = m.position_encoder.psne_layer
    m.load_state_dict(own, strict=True)
    return m, inf


@need_ref
def test_forward_and_loss_match_live_reference():
    sd = synth.make_caption_params(3)
    sd['caption_generator.position_encoder.psne_layer'] = CO.positions(35, 768)
    m, _ = _live(sd)
    memory, ids, mask = case(1)
    embs = CO.embed_ids(sd, ids)
    with torch.no_grad():
        outs, logits = m(tgt=embs[:, :-1, :], memory=memory, tgt_key_padding_mask=torch.logical_not(mask.bool()[:, :-1]))
        o_outs, o_logits = CO.forward(sd, CFG, embs[:, :-1, :], memory, torch.logical_not(mask.bool()[:, :-1]))
    assert torch.equal(m.position_encoder.psne_layer, CO.positions(35, 768))
    for a, b in zip(outs, o_outs):
        assert float((a - b).abs().max()) < 1e-5
    assert float((logits - o_logits).abs().max()) < 1e-4 * float(logits.abs().max())
    want = 2.0 * torch.nn.functional.cross_entropy(logits.flatten(0, 1), ids[:, 1:].flatten(), reduction='none', ignore_index=0).mean()
    got = CO.caption_loss(sd, CFG, memory, ids, embs, mask)
    assert abs(float(got) - float(want)) < 1e-5 * abs(float(want))


@need_ref
def test_beam_search_matches_live_reference():
    sd = synth.make_caption_params(5)
    sd['caption_generator.position_encoder.psne_layer'] = CO.positions(35, 768)
    m, inf = _live(sd)
    memory = case(2, B=1)[0]

    class _Tok:
        def decode(self, ids):
            return '[' + ' '.join(str(i) for i in ids) + ']'
    inf.transformers = types.SimpleNamespace(BertTokenizer=types.SimpleNamespace(from_pretrained=lambda name: _Tok()))
    be = types.SimpleNamespace(
        word_embeddings=lambda ids: torch.nn.functional.embedding(ids, sd['bert_embeddings.word_embeddings.weight']),
        LayerNorm=lambda e: torch.nn.functional.layer_norm(e, (768,), sd['bert_embeddings.LayerNorm.weight'],
                                                           sd['bert_embeddings.LayerNorm.bias'], 1e-12))
    model = types.SimpleNamespace(caption_generator=m, bert_embeddings=be)
    with torch.no_grad():
        text = inf.beam_search(model, memory, BOS, EOS, max_len=35, beam_width=7)
    best, finished = CO.beam_search(sd, CFG, memory, BOS, EOS)
    assert best is not None and text == ' '.join(str(i) for i in best)


def test_golden_fixture():
    z = np.load(os.path.join(HERE, 'golden', 'caption.npz'))
    sd = synth.make_caption_params(int(z['seed']))
    sd['caption_generator.position_encoder.psne_layer'] = CO.positions(35, 768)
    memory, ids, mask = case(int(z['case_seed']))
    embs = CO.embed_ids(sd, ids)
    logits = CO.forward(sd, CFG, embs[:, :-1, :], memory, torch.logical_not(mask.bool()[:, :-1]))[1]
    assert float((logits[:, :, :64] - torch.from_numpy(z['logits_head'])).abs().max()) < 2e-4 * float(np.abs(z['logits_head']).max())
    assert abs(float(CO.caption_loss(sd, CFG, memory, ids, embs, mask)) - float(z['loss'])) < 1e-5 * abs(float(z['loss']))
    best, _ = CO.beam_search(sd, CFG, memory[:1], BOS, EOS)
    assert best == list(z['beam_ids'])


def test_product_state_dict_keys_and_bit_packing():
    from cgg_b200.caption import CaptionTransformerB200, _pack_bits
    m = CaptionTransformerB200(**CFG)
    want = {k[len('caption_generator.'):] for k in synth.make_caption_params(0) if k.startswith('caption_generator.')}
    assert set(m.state_dict()) == want | {'position_encoder.psne_layer'}
    assert torch.equal(m.position_encoder.psne_layer, CO.positions(35, 768))
    g = torch.Generator().manual_seed(0)
    mask = torch.rand((2, 5, 70), generator=g) > 0.5
    words = _pack_bits(mask)
    for b in range(2):
        for q in range(5):
            for k in range(70):
                assert bool((int(words[b, q, k // 32]) >> (k % 32)) & 1) == bool(mask[b, q, k])
