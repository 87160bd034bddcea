"""GPU tests of the caption generator (cgg_b200/caption.py) against the oracle
(oracle/caption_oracle.py, pinned on the live reference modules) and the reference-generated fixture
tests/golden/caption.npz: forward of every block, the caption-generation loss with the gradient of every parameter and of
the query embeddings (fp32 FMA mode, and the tf32 tensor-core mode within its own tolerance), the beam search's tokens."""
import os
import numpy as np
import pytest
import torch

from oracle import caption_oracle as CO
from cgg_b200 import synth
from cgg_b200.head import Mask2FormerHeadOpenB200
from cgg_b200.caption import caption_generation_loss, beam_search
from test_caption_cpu import case, CFG, BOS, EOS

pytestmark = pytest.mark.gpu
DEV = 'cuda'
HERE = os.path.dirname(os.path.abspath(__file__))


def _head(seed, train_precision='fp32'):
    sd = synth.make_caption_params(seed)
    head = Mask2FormerHeadOpenB200(num_things_classes=48, num_stuff_classes=0, num_queries=12, use_class_emb=True,
                                   use_caption=True, use_caption_generation=True, bert_vocab_size=CFG['nb_tokens'],
                                   caption_generator=dict(type='CaptionTransformer', **CFG), train_precision=train_precision,
                                   loss_caption_generation=dict(type='CrossEntropyLoss', ignore_index=0, loss_weight=2.0))
    missing = head.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all(not k.startswith(('caption_generator.t', 'caption_generator.g', 'bert'))
                                               for k in missing.missing_keys)
    sd['caption_generator.position_encoder.psne_layer'] = CO.positions(35, 768)
    return head.to(DEV), sd


@pytest.mark.parametrize('train_precision', ['fp32', 'tf32'])
def test_forward_loss_and_gradients_match_oracle(train_precision):
    head, sd = _head(3, train_precision)
    head.train()
    memory, ids, mask = case(1)
    embs = CO.embed_ids(sd, ids)
    sd_o = {k: v.clone().requires_grad_(k.startswith('caption_generator.') and 'psne' not in k) for k, v in sd.items()}
    mem_o = memory.clone().requires_grad_(True)
    want_outs, want_logits = CO.forward(sd_o, CFG, embs[:, :-1, :], mem_o, torch.logical_not(mask.bool()[:, :-1]))
    want = CO.caption_loss(sd_o, CFG, mem_o, ids, embs, mask)
    want.backward()
    mem_d = memory.to(DEV).requires_grad_(True)
    outs, logits = head.caption_generator(tgt=embs[:, :-1, :].to(DEV), memory=mem_d,
                                          tgt_key_padding_mask=torch.logical_not(mask.bool()[:, :-1]).to(DEV))
    tol = dict(fp32=(2e-5, 1e-5, 2e-4), tf32=(1e-2, 5e-3, 5e-2))[train_precision]        # forward, loss, gradients
    for a, b in zip(outs, want_outs):
        assert float((a.detach().cpu() - b.detach()).abs().max()) < tol[0] * float(b.abs().max())
    assert float((logits.detach().cpu() - want_logits.detach()).abs().max()) < tol[0] * float(want_logits.abs().max())
    got = caption_generation_loss(head, mem_d, list(ids.to(DEV)), list(embs.to(DEV)), list(mask.to(DEV)), None, 2.0)
    assert abs(float(got) - float(want)) < tol[1] * abs(float(want))
    got.backward()
    worst = ('', 0.0)
    # (a key-projection bias shifts every logit of a query equally: its true gradient is 0 and the oracle's is rounding
    # noise -- gradients are compared relative to max(their own scale, 1e-3 of the largest gradient in the module))
    gmax = max(float(sd_o['caption_generator.' + k].grad.abs().max()) for k, _ in head.caption_generator.named_parameters())
    for k, p in head.caption_generator.named_parameters():
        w = sd_o['caption_generator.' + k].grad
        assert p.grad is not None, k
        if train_precision == 'tf32':     # ReLU gates flip under operand rounding (see test_gpu_train.py): relative L2 error
            err = float((p.grad.cpu() - w).norm()) / max(float(w.norm()), 1e-3 * gmax * w.numel() ** 0.5)
        else:
            err = float((p.grad.cpu() - w).abs().max()) / max(float(w.abs().max()), 1e-3 * gmax)
        worst = max(worst, (k, err), key=lambda t: t[1])
    assert worst[1] < tol[2], worst
    err = float((mem_d.grad.cpu() - mem_o.grad).abs().max()) / float(mem_o.grad.abs().max())
    assert err < tol[2], err
    print('[%s] worst caption-generator gradient error %s %.2e; memory %.2e' % (train_precision, worst[0], worst[1], err))


def test_beam_search_and_fixture():
    z = np.load(os.path.join(HERE, 'golden', 'caption.npz'))
    head, sd = _head(int(z['seed']))
    head.eval()
    memory, ids, mask = case(int(z['case_seed']))
    embs = CO.embed_ids(sd, ids)
    with torch.no_grad():
        logits = head.caption_generator(tgt=embs[:, :-1, :].to(DEV), memory=memory.to(DEV),
                                        tgt_key_padding_mask=torch.logical_not(mask.bool()[:, :-1]).to(DEV))[1]
        loss = caption_generation_loss(head, memory.to(DEV), list(ids.to(DEV)), list(embs.to(DEV)), list(mask.to(DEV)), None, 2.0)
    assert float((logits[:, :, :64].cpu() - torch.from_numpy(z['logits_head'])).abs().max()) < 2e-4 * float(np.abs(z['logits_head']).max())
    assert abs(float(loss) - float(z['loss'])) < 1e-5 * abs(float(z['loss']))
    res = beam_search(head, memory[:1].to(DEV), BOS, EOS, max_len=35, beam_width=7)
    assert res['ids'] == [int(t) for t in z['beam_ids']]            # the reference's own sentence
    best, finished = CO.beam_search(sd, CFG, memory[:1], BOS, EOS)
    assert [f[0] for f in res['finished']] == [f[0] for f in finished]
    for a, b in zip(res['finished'], finished):
        assert abs(a[1] - b[1]) < 1e-4 * abs(b[1])


def test_simple_test_with_caption_and_forward_train_loss_key():
    """simple_test(with_caption=True) (head.py:966-970) returns the beam search's sentence (token ids without a tokenizer);
    forward_train adds loss_caption_generation for all 10 head calls (loss_single :550-583)."""
    from test_gpu_post import _StubPixelDecoder
    Q, B = 12, 2
    sdh = synth.make_params(seed=6, num_queries=Q, perturb=True)
    sdc = synth.make_caption_params(5)
    head = Mask2FormerHeadOpenB200(num_things_classes=48, num_stuff_classes=0, num_queries=Q, use_class_emb=True,
                                   use_caption=True, use_caption_generation=True, bert_vocab_size=CFG['nb_tokens'],
                                   caption_generator=dict(type='CaptionTransformer', **CFG), pixel_decoder=_StubPixelDecoder(),
                                   loss_grounding=dict(type='GroundingLoss', loss_weight=2.0),
                                   loss_caption_generation=dict(type='CrossEntropyLoss', ignore_index=0, loss_weight=2.0))
    head.load_state_dict({**sdh, **sdc}, strict=False)
    head = head.to(DEV)
    mf, mems = synth.make_inputs(8, B, 96, 128)
    feats = [mf.to(DEV)] + [m.to(DEV) for m in mems]
    _, ids, mask = case(3)
    head.train()
    losses = head.forward_train(feats, [dict() for _ in range(B)], None, None, None, None, list(ids.to(DEV)), list(mask.to(DEV)),
                                list(ids.to(DEV)), list(mask.to(DEV)))
    assert {'loss_caption_generation'} | {'d%d.loss_caption_generation' % j for j in range(9)} <= set(losses)
    sum(losses.values()).backward()
    assert head.caption_generator.generator.weight.grad is not None and head.v2l_transform.weight.grad is not None
    head.eval()
    with torch.no_grad():
        out = head.simple_test([f[:1] for f in feats], [dict(batch_input_shape=(96, 128))], with_caption=True)
    assert out[3] is None or (isinstance(out[3], list) and out[3][0] == BOS)
