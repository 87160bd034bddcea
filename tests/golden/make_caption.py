"""Generates tests/golden/caption.npz from the UNMODIFIED reference (CaptionTransformer forward, the caption-generation
cross entropy, beam_search with the tokenizer download stubbed) on seeded synthetic weights (cgg_b200/synth.py) -- the
fixture that travels to the GPU box.   python tests/golden/make_caption.py"""
import os
import sys
import types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import caption_oracle as CO      # noqa: E402  (only for the embedding lookup helper)
from cgg_b200 import synth                   # noqa: E402
import test_caption_cpu as T                 # noqa: E402


def main():
    seed, case_seed = 11, 4
    sd = synth.make_caption_params(seed)
    sd['caption_generator.position_encoder.psne_layer'] = CO.positions(35, 768)
    m, inf = T._live(sd)
    memory, ids, mask = T.case(case_seed)
    embs = CO.embed_ids(sd, ids)
    with torch.no_grad():
        logits = m(tgt=embs[:, :-1, :], memory=memory, tgt_key_padding_mask=torch.logical_not(mask.bool()[:, :-1]))[1]
    loss = 2.0 * torch.nn.functional.cross_entropy(logits.flatten(0, 1), ids[:, 1:].flatten(), reduction='none', ignore_index=0).mean()

    class _Tok:
        def decode(self, ids_):
            return '[' + ' '.join(str(i) for i in ids_) + ']'
    inf.transformers = types.SimpleNamespace(BertTokenizer=types.SimpleNamespace(from_pretrained=lambda name: _Tok()))
    be = types.SimpleNamespace(
        word_embeddings=lambda i: torch.nn.functional.embedding(i, sd['bert_embeddings.word_embeddings.weight']),
        LayerNorm=lambda e: torch.nn.functional.layer_norm(e, (768,), sd['bert_embeddings.LayerNorm.weight'],
                                                           sd['bert_embeddings.LayerNorm.bias'], 1e-12))
    with torch.no_grad():
        text = inf.beam_search(types.SimpleNamespace(caption_generator=m, bert_embeddings=be), memory[:1], T.BOS, T.EOS,
                               max_len=35, beam_width=7)
    beam = [int(t) for t in text.split()]
    np.savez_compressed(os.path.join(HERE, 'caption.npz'), seed=seed, case_seed=case_seed, logits_head=logits[:, :, :64].numpy(),
                        loss=float(loss), beam_ids=np.array(beam))
    print(float(loss), beam)


if __name__ == '__main__':
    main()
