"""GPU tests of the training step (BASELINE.json configs[3]; reference forward_train -> loss -> backward,
mask2former_head.py:851-921 / :393-462): the autograd-connected forward of cgg_b200/train.py, whose forward AND backward
nodes are kernels of the C-ABI library, against the oracle's autograd (plain torch ops on the CPU, fp32) -- outputs and
the gradient of EVERY state_dict key, of the mask features and of the three memories."""
import pytest
import torch

from oracle import cgg_oracle as O
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
from cgg_b200.grounding import grounding_loss

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _loss_from_outputs(cls, emb, mask, probes, cap, cap_mask, gl):
    """A scalar that touches every output of every head call: fixed random probes for cls / emb / mask logits (a stand-in
    for the mask BCE/dice and class losses, whose matching step is outside the path) plus the caption-grounding loss of
    every head call with the reference's weight 2.0 (configs/openset_panoptic/coco_panoptic_p20.py:123-125)."""
    total = 0.0
    for j in range(len(cls)):
        total = total + (cls[j] * probes['cls'][j]).sum() + (emb[j] * probes['emb'][j]).sum() * 0.1 \
            + (mask[j] * probes['mask'][j]).sum() * 0.01
        total = total + gl(emb[j], cap, cap_mask)
    return total


def _setup(Q, B, H, W, ncls1=49, seed=41):
    sd = synth.make_params(seed=seed, num_queries=Q, perturb=True, num_classes_p1=ncls1)
    mf, mems = synth.make_inputs(seed, B, H, W)
    g = torch.Generator().manual_seed(seed)
    H4, W4 = mf.shape[-2:]
    probes = dict(cls=[torch.randn((B, Q, ncls1), generator=g) for _ in range(10)],
                  emb=[torch.randn((B, Q, 768), generator=g) for _ in range(10)],
                  mask=[torch.randn((B, Q, H4, W4), generator=g) for _ in range(10)])
    ids, cap_mask, table, lw, lb = synth.make_captions(seed, B, vocab=2000)
    cap = O.noun_embeddings(table, lw, lb, ids)
    return sd, mf, mems, probes, cap, cap_mask


# tolerances (relative to each tensor's own max-abs): fp32 FMA mode is the parity mode; in tf32 mode every contraction
# rounds its operands to 10-bit mantissas (what torch.backends.cuda.matmul.allow_tf32 does to the reference's own step)
TOL = {'fp32': dict(fwd=5e-4, loss=1e-4, grad=2e-3), 'tf32': dict(fwd=5e-3, loss=2e-3, grad=5e-2)}


@pytest.mark.parametrize('train_precision', ['fp32', 'tf32'])
@pytest.mark.parametrize('Q,B,H,W,ncls1', [(24, 2, 128, 160, 49), (40, 1, 160, 128, 118)])
def test_training_step_gradients_match_oracle_autograd(Q, B, H, W, ncls1, train_precision):
    tol = TOL[train_precision]
    sd, mf, mems, probes, cap, cap_mask = _setup(Q, B, H, W, ncls1)
    # ---- oracle: plain torch autograd on the CPU
    sd_o = {k: v.clone().requires_grad_(k != 'class_embs') for k, v in sd.items()}
    mf_o = mf.clone().requires_grad_(True)
    mems_o = [m.clone().requires_grad_(True) for m in mems]
    ref = O.decoder_forward(sd_o, mf_o, mems_o)
    loss_o = _loss_from_outputs(ref['cls'], ref['emb'], ref['mask'], probes, cap, cap_mask,
                                lambda e, c, m: O.grounding_loss(e, c, m, 10.0, 2.0))
    loss_o.backward()
    # ---- ours
    head = build_head_from_state_dict(sd, Q, ncls1, 'fp32', DEV, train_precision=train_precision).train()
    mf_d = mf.to(DEV).requires_grad_(True)
    mems_d = [m.to(DEV).requires_grad_(True) for m in mems]
    # free-running first: the attention masks this forward derives from its own logits agree with the oracle's
    from cgg_b200.train import decoder_forward_train
    with torch.no_grad():
        dbg = head.decoder_forward(mf_d.detach(), [m.detach() for m in mems_d], return_debug=True)[3]
    want_bits = [O.pack_mask_bits(ref['masked'][j].detach()) for j in range(9)]
    nbits = sum(w.numel() * 32 for w in want_bits)
    nflip = sum(int(torch.tensor([bin(int(x) & 0xffffffff).count('1') for x in (dbg['bitmaps'][j].cpu() ^ want_bits[j]).flatten()]).sum())
                for j in range(9))
    assert nflip <= 2e-4 * nbits, (nflip, nbits)
    # the gradient comparison itself runs on the ORACLE's attention masks (they are detached constants of the graph,
    # head.py:759), so a threshold-band bit of a 20-key level cannot derail it
    forced = [(want_bits[j].to(DEV), ref['masked'][j].detach().all(-1).to(torch.uint8).to(DEV)) for j in range(9)]
    cls, emb, mask = decoder_forward_train(head, mf_d, mems_d, forced_attn_masks=forced)
    assert cls[0].requires_grad and mask[9].requires_grad
    probes_d = {k: [t.to(DEV) for t in v] for k, v in probes.items()}
    loss = _loss_from_outputs(cls, emb, mask, probes_d, cap.to(DEV), cap_mask.to(DEV),
                              lambda e, c, m: grounding_loss(e, c, m, 10.0, 2.0))
    loss.backward()
    torch.cuda.synchronize()
    # forward values
    for j in range(10):
        assert float((mask[j].detach().cpu() - ref['mask'][j].detach()).abs().max()) < tol['fwd'] * float(ref['mask'][j].abs().max())
        assert float((emb[j].detach().cpu() - ref['emb'][j].detach()).abs().max()) < tol['fwd'] * float(ref['emb'][j].abs().max())
    assert abs(float(loss) - float(loss_o)) < tol['loss'] * abs(float(loss_o))
    # gradient of every parameter (state_dict key), relative to the gradient's own scale
    worst = ('', 0.0)
    named = dict(head.named_parameters())
    assert set(named) == {k for k in sd if k != 'class_embs'}
    for k, p in named.items():
        assert p.grad is not None, 'no gradient reached %s' % k
        want = sd_o[k].grad
        err = float((p.grad.cpu() - want).abs().max()) / (float(want.abs().max()) + 1e-12)
        if err > worst[1]:
            worst = (k, err)
        assert err < tol['grad'], (k, err)
    # gradients of the path's inputs (they flow on into the pixel decoder in the real model)
    for got, want in [(mf_d.grad, mf_o.grad)] + [(a.grad, b_.grad) for a, b_ in zip(mems_d, mems_o)]:
        err = float((got.cpu() - want).abs().max()) / float(want.abs().max())
        assert err < tol['grad'], err
    print('[%s] worst parameter gradient error: %s %.2e' % ((train_precision,) + worst))
    if train_precision == 'tf32':
        # calibration: the same step through plain torch CUDA ops with TF32 matmuls (what the reference runs with
        # allow_tf32) against the same fp32 oracle -- the tf32 mode must not be further from fp32 than that by much
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            sd_t = {k: v.clone().to(DEV).requires_grad_(k != 'class_embs') for k, v in sd.items()}
            ref_t = O.decoder_forward(sd_t, mf.to(DEV), [m.to(DEV) for m in mems],
                                      forced_masked=[r.detach().to(DEV) for r in ref['masked']])
            loss_t = _loss_from_outputs(ref_t['cls'], ref_t['emb'], ref_t['mask'], probes_d, cap.to(DEV), cap_mask.to(DEV),
                                        lambda e, c, m: O.grounding_loss(e, c, m, 10.0, 2.0))
            loss_t.backward()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
        worst_t = max(float((sd_t[k].grad.cpu() - sd_o[k].grad).abs().max()) / (float(sd_o[k].grad.abs().max()) + 1e-12)
                      for k in named)
        print('[torch TF32 matmuls] worst parameter gradient error: %.2e' % worst_t)


def test_forward_dispatch_inference_vs_training():
    """`forward` records a graph only when gradients are being recorded; under no_grad it is the inference path."""
    Q, B = 16, 1
    sd = synth.make_params(seed=3, num_queries=Q, perturb=True)
    mf, mems = synth.make_inputs(3, B, 96, 128)
    head = build_head_from_state_dict(sd, Q, 49, 'fp32', DEV)
    mfd, memd = mf.to(DEV), [m.to(DEV) for m in mems]
    with torch.no_grad():
        a = head.decoder_forward_auto(mfd, memd)
    b = head.decoder_forward_auto(mfd, memd)
    assert not a[2][9].requires_grad and b[2][9].requires_grad
    for j in range(10):
        assert float((a[2][j] - b[2][j].detach()).abs().max()) < 2e-4
        assert float((a[1][j] - b[1][j].detach()).abs().max()) < 2e-4
