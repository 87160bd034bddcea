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


# fp32 (FMA) mode is the parity mode: tolerances relative to each tensor's own max-abs.  In tf32 mode every contraction
# rounds its operands to 10-bit mantissas -- what torch.backends.cuda.matmul.allow_tf32 does to the reference's own step --
# and the gradients of a randomly initialised 9-layer ReLU network are sensitive to that (a pre-activation near zero
# changes its gate): the forward outputs are held to BASELINE.json's 1e-2-of-range bar, the gradients are calibrated
# against the same step run through plain torch CUDA ops with TF32 matmuls (same oracle code, same masks).
TOL = {'fp32': dict(fwd=5e-4, loss=1e-4, grad=2e-3), 'tf32': dict(fwd=1e-2, loss=None, grad=None)}
TF32_VS_TORCH_TF32 = 4.0        # tf32 mode may be at most this many times further from fp32 than torch's TF32 step
                                # (tcgen05 truncates the operands to tf32, cuBLAS rounds them to nearest; the attention
                                # products are tf32 too, torch's fused attention stays fp32)


def _rel(a, b):
    return float((a.detach().cpu() - b.detach().cpu()).abs().max()) / (float(b.detach().abs().max()) + 1e-12)


def _l2(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).norm() / (b.detach().double().norm() + 1e-30))


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
        assert _rel(mask[j], ref['mask'][j]) < tol['fwd']
        assert _rel(emb[j], ref['emb'][j]) < tol['fwd']
        assert _rel(cls[j], ref['cls'][j]) < tol['fwd']
    named = dict(head.named_parameters())
    assert set(named) == {k for k in sd if k != 'class_embs'}
    for k, p in named.items():
        assert p.grad is not None, 'no gradient reached %s' % k
    inputs = [(mf_d.grad, mf_o.grad)] + [(a.grad, b_.grad) for a, b_ in zip(mems_d, mems_o)]
    worst = max((_rel(p.grad, sd_o[k].grad), k) for k, p in named.items())
    worst_l2 = max((_l2(p.grad, sd_o[k].grad), k) for k, p in named.items())
    worst_in = max(_rel(g, w) for g, w in inputs)
    print('[%s] worst parameter gradient error: max-abs %.2e (%s), L2 %.2e (%s); inputs %.2e'
          % (train_precision, worst[0], worst[1], worst_l2[0], worst_l2[1], worst_in))
    if train_precision == 'fp32':
        assert abs(float(loss) - float(loss_o)) < tol['loss'] * abs(float(loss_o))
        assert worst[0] < tol['grad'], worst          # gradient of every parameter (state_dict key)
        assert worst_in < tol['grad'], worst_in       # and of the path's inputs (they flow on into the pixel decoder)
        return
    # ---- tf32: calibrate on the same step through plain torch CUDA ops with TF32 matmuls
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        sd_t = {k: v.clone().to(DEV).requires_grad_(k != 'class_embs') for k, v in sd.items()}
        mf_t = mf.to(DEV).requires_grad_(True)
        mems_t = [m.to(DEV).requires_grad_(True) for m in mems]
        ref_t = O.decoder_forward(sd_t, mf_t, mems_t, forced_masked=[r.detach().to(DEV) for r in ref['masked']])
        loss_t = _loss_from_outputs(ref_t['cls'], ref_t['emb'], ref_t['mask'], probes_d, cap.to(DEV), cap_mask.to(DEV),
                                    lambda e, c, m: O.grounding_loss(e, c, m, 10.0, 2.0))
        loss_t.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    t_worst = max(_rel(sd_t[k].grad, sd_o[k].grad) for k in named)
    t_l2 = max(_l2(sd_t[k].grad, sd_o[k].grad) for k in named)
    t_in = max(_rel(g, w) for g, w in [(mf_t.grad, mf_o.grad)] + [(a.grad, b_.grad) for a, b_ in zip(mems_t, mems_o)])
    t_loss = abs(float(loss_t) - float(loss_o))
    print('[torch TF32 matmuls] worst parameter gradient error: max-abs %.2e, L2 %.2e; inputs %.2e; loss %.2e (ours %.2e)'
          % (t_worst, t_l2, t_in, t_loss, abs(float(loss) - float(loss_o))))
    assert worst[0] < TF32_VS_TORCH_TF32 * t_worst, (worst, t_worst)
    assert worst_l2[0] < TF32_VS_TORCH_TF32 * t_l2, (worst_l2, t_l2)
    assert worst_in < TF32_VS_TORCH_TF32 * t_in, (worst_in, t_in)
    assert abs(float(loss) - float(loss_o)) < 2 * TF32_VS_TORCH_TF32 * t_loss + 1e-3 * abs(float(loss_o))


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


def test_graphed_step_replays_the_eager_gradients():
    """GraphedStep: forward + loss + backward captured once as a CUDA graph; a replay on new input values gives the
    gradients of an eager step on those values (up to the summation order of the bias-gradient atomics)."""
    from cgg_b200.train import GraphedStep, GradReducer
    Q, B, ncls1 = 24, 2, 49
    sd, mf, mems, probes, cap, cap_mask = _setup(Q, B, 128, 160, ncls1)
    head = build_head_from_state_dict(sd, Q, ncls1, 'fp32', DEV, train_precision='tf32').train()
    probes_d = {k: [t.to(DEV) for t in v] for k, v in probes.items()}
    cap_d, cm_d = cap.to(DEV), cap_mask.to(DEV)
    s_mf, s_mems = mf.to(DEV).clone(), [m.to(DEV).clone() for m in mems]

    def step_fn():
        cls, emb, mask = head.decoder_forward_auto(s_mf, s_mems)
        return _loss_from_outputs(cls, emb, mask, probes_d, cap_d, cm_d, lambda e, c, m: grounding_loss(e, c, m, 10.0, 2.0))

    for use_reducer in (False, True):
        red = GradReducer(head.parameters(), bucket_mb=4.0) if use_reducer else None
        gs = GraphedStep(step_fn, head.parameters(), reducer=red)
        # new input values, written into the static tensors
        mf2, mems2 = synth.make_inputs(77, B, 128, 160)
        s_mf.copy_(mf2.to(DEV))
        for a, b_ in zip(s_mems, mems2):
            a.copy_(b_.to(DEV))
        loss_g = float(gs.replay())
        torch.cuda.synchronize()
        got = {k: p.grad.clone() for k, p in head.named_parameters()}
        if red is not None:
            red.remove()
        for p in head.parameters():
            p.grad = None
        loss_e = step_fn()
        loss_e.backward()
        torch.cuda.synchronize()
        assert abs(loss_g - float(loss_e)) <= 1e-6 * abs(float(loss_e))
        for k, p in head.named_parameters():
            assert float((got[k] - p.grad).abs().max()) <= 1e-5 * float(p.grad.abs().max()) + 1e-12, (use_reducer, k)
        for p in head.parameters():
            p.grad = None
        del loss_e, gs        # (an autograd graph built on the default stream must not outlive into the next capture:
        #                        its AccumulateGrad nodes are bound to that stream)


@pytest.mark.parametrize('train_precision', ['fp32', 'tf32'])
def test_fused_weight_gradients_equal_the_autograd_path(train_precision):
    """fused_wgrad: dW / db of the linear layers accumulated into pre-existing .grad tensors by the GEMM epilogue on a side
    stream (train._WgradSide) -- same gradients as the autograd path, and they ADD to what .grad already holds."""
    from cgg_b200.train import decoder_forward_train
    Q, B, ncls1 = 24, 2, 49
    sd, mf, mems, probes, cap, cap_mask = _setup(Q, B, 128, 160, ncls1)
    probes_d = {k: [t.to(DEV) for t in v] for k, v in probes.items()}

    def run(fused):
        head = build_head_from_state_dict(sd, Q, ncls1, 'fp32', DEV, train_precision=train_precision,
                                          fused_wgrad='always' if fused else False).train()
        if fused:
            for p in head.parameters():
                p.grad = torch.full_like(p, 0.25)          # pre-existing content must be kept (accumulation)
        cls, emb, mask = decoder_forward_train(head, mf.to(DEV), [m.to(DEV) for m in mems])
        loss = _loss_from_outputs(cls, emb, mask, probes_d, cap.to(DEV), cap_mask.to(DEV),
                                  lambda e, c, m: grounding_loss(e, c, m, 10.0, 2.0))
        loss.backward()
        torch.cuda.synchronize()
        return {k: (p.grad - 0.25 if fused else p.grad).clone() for k, p in head.named_parameters()}

    a, b_ = run(False), run(True)
    for k in a:
        scale = float(a[k].abs().max()) + 1e-12
        assert float((a[k] - b_[k]).abs().max()) <= 2e-5 * scale + 1e-6, (k, float((a[k] - b_[k]).abs().max()), scale)
