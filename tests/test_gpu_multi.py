"""2-rank NCCL tests (need >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py`; skipped on a
single-GPU box): the two real exchange steps of the training config --
  * the batched caption / prediction all-gather before the grounding loss (mask2former_head.py:650-684) on the CUDA
    `all_gather_into_tensor` branch, with OUR grounding-loss kernels on the gathered operands, against the single-process
    oracle loss and its gradient;
  * the bucketed gradient all-reduce of a data-parallel training step (open_set/apis/train.py:156-161): the reduced
    gradients must equal the gradient of the loss averaged over the global batch, computed by the oracle's autograd."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
need2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))


def _np(x):
    """Results cross the queue BY VALUE (numpy): a torch tensor would be handed over through the sender's fd-sharing
    socket, which is gone if the worker exits before the parent unpickles."""
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    if isinstance(x, dict):
        return {k: _np(v) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return tuple(_np(v) for v in x)
    return x


def _pt(x):
    import numpy as np
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x)
    if isinstance(x, dict):
        return {k: _pt(v) for k, v in x.items()}
    if isinstance(x, tuple):
        return tuple(_pt(v) for v in x)
    return x


def _join(procs, timeout=90):
    """Waits for the workers; a worker that is still alive is killed (never left behind on the box) and reported."""
    codes = []
    for p in procs:
        p.join(timeout=timeout)
        if p.is_alive():
            p.kill()
            p.join(timeout=10)
            codes.append('hung')
        else:
            codes.append(p.exitcode)
    assert codes == [0] * len(procs), codes


def _gather_data(world, L=3, B=2, Q=20, T=35, D=768):
    g = torch.Generator().manual_seed(5)
    pred = torch.randn((L, world * B, Q, D), generator=g)
    cap = torch.randn((world * B, T, D), generator=g) * 0.85
    mask = torch.zeros((world * B, T), dtype=torch.long)
    for b, n in enumerate([3, 35, 0, 7][:world * B]):
        mask[b, :n] = 1
    return pred, cap, mask


def _gather_worker(rank, world, port, q):
    _init(rank, world, port)
    from cgg_b200.grounding import gather_captions_and_preds, grounding_loss
    dev = torch.device('cuda', rank)
    pred_all, cap_all, mask_all = _gather_data(world)
    B = 2
    sl = slice(rank * B, (rank + 1) * B)
    pred = pred_all[:, sl].clone().to(dev).requires_grad_(True)
    embs, mask, preds = gather_captions_and_preds(list(cap_all[sl].to(dev)), list(mask_all[sl].to(dev)), pred)
    loss = sum(grounding_loss(preds[l], embs, mask, 10.0, 2.0) for l in range(preds.shape[0]))
    loss.backward()
    q.put(_np((rank, embs, mask, preds, float(loss), pred.grad)))
    dist.barrier()
    dist.destroy_process_group()


@need2
def test_nccl_caption_gather_and_grounding_loss():
    sys.path.insert(0, ROOT)
    from oracle import cgg_oracle as O
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted((_pt(q.get(timeout=150)) for _ in range(2)), key=lambda r: r[0])
    finally:
        _join(procs)
    pred_all, cap_all, mask_all = _gather_data(2)
    full = pred_all.clone().requires_grad_(True)
    want = sum(O.grounding_loss(full[l], cap_all, mask_all, 10.0, 2.0) for l in range(full.shape[0]))
    want.backward()
    for rank, embs, mask, preds, loss, grad in res:
        assert torch.equal(embs, cap_all) and torch.equal(mask, mask_all)
        assert torch.equal(preds, pred_all)
        assert abs(loss - float(want)) < 2e-4 * max(1.0, abs(float(want))), (loss, float(want))
        want_g = full.grad[:, rank * 2:(rank + 1) * 2]      # the similarity GEMMs run at split-bf16 precision (~1e-5 of max)
        assert float((grad - want_g).abs().max()) < 2e-4 * float(want_g.abs().max())


def _train_worker(rank, world, port, q):
    _init(rank, world, port)
    from cgg_b200 import synth
    from cgg_b200.head import build_head_from_state_dict
    from cgg_b200.train import GradReducer
    dev = torch.device('cuda', rank)
    Q, B = 16, 1
    sd = synth.make_params(seed=51, num_queries=Q, perturb=True)
    head = build_head_from_state_dict(sd, Q, 49, 'fp32', dev).train()
    mf, mems = synth.make_inputs(60 + rank, B, 96, 128)
    red = GradReducer(head.parameters(), bucket_mb=4.0)
    assert len(red.buckets) > 1
    cls, emb, mask = head.decoder_forward_auto(mf.to(dev), [m.to(dev) for m in mems])
    loss = sum((c ** 2).mean() + (e ** 2).mean() + (m ** 2).mean() for c, e, m in zip(cls, emb, mask))
    loss.backward()
    red.finish()
    torch.cuda.synchronize()
    q.put(_np((rank, {k: p.grad for k, p in head.named_parameters()}, red.exposed())))
    dist.barrier()
    dist.destroy_process_group()


def _graph_worker(rank, world, port, q):
    _init(rank, world, port)
    from cgg_b200 import synth
    from cgg_b200.head import build_head_from_state_dict
    from cgg_b200.train import GradReducer, GraphedStep
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    Q, B = 16, 1
    sd = synth.make_params(seed=51, num_queries=Q, perturb=True)
    head = build_head_from_state_dict(sd, Q, 49, 'fp32', dev, train_precision='tf32').train()
    mf, mems = synth.make_inputs(60 + rank, B, 96, 128)
    s_mf, s_mems = mf.to(dev), [m.to(dev) for m in mems]

    def step_fn():
        cls, emb, mask = head.decoder_forward_auto(s_mf, s_mems)
        return sum((c ** 2).mean() + (e ** 2).mean() + (m ** 2).mean() for c, e, m in zip(cls, emb, mask))

    red = GradReducer(head.parameters(), bucket_mb=4.0)
    gs = GraphedStep(step_fn, head.parameters(), reducer=red)         # the bucketed all-reduces are part of the graph
    gs.replay()
    gs.replay()
    torch.cuda.synchronize()
    graphed = {k: p.grad.clone() for k, p in head.named_parameters()}
    red.zero()                                                         # the same step eagerly, same reducer
    step_fn().backward()
    red.finish()
    torch.cuda.synchronize()
    q.put(_np((rank, {k: (graphed[k], p.grad) for k, p in head.named_parameters()})))
    red.remove()
    del gs, red                                                        # the graph holds captured NCCL work: drop it first
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@need2
def test_graphed_training_step_with_nccl_allreduce_inside_the_graph():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_graph_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted((_pt(q.get(timeout=150)) for _ in range(2)), key=lambda r: r[0])
    finally:
        _join(procs)
    for k in res[0][1]:
        g0, e0 = res[0][1][k]
        g1, e1 = res[1][1][k]
        assert torch.equal(g0, g1), k                                   # both ranks hold the averaged gradient
        assert float((g0 - e0).abs().max()) <= 1e-6 * float(e0.abs().max()) + 1e-12, k     # graph replay == eager step


@need2
def test_nccl_gradient_allreduce_matches_global_batch_gradient():
    sys.path.insert(0, ROOT)
    from oracle import cgg_oracle as O
    from cgg_b200 import synth
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 34500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted((_pt(q.get(timeout=150)) for _ in range(2)), key=lambda r: r[0])
    finally:
        _join(procs)
    sd = synth.make_params(seed=51, num_queries=16, perturb=True)
    sd_o = {k: v.clone().requires_grad_(k != 'class_embs') for k, v in sd.items()}
    total = 0.0
    for rank in range(2):
        mf, mems = synth.make_inputs(60 + rank, 1, 96, 128)
        ref = O.decoder_forward(sd_o, mf, mems)
        total = total + sum((c ** 2).mean() + (e ** 2).mean() + (m ** 2).mean()
                            for c, e, m in zip(ref['cls'], ref['emb'], ref['mask']))
    (total / 2).backward()
    for rank, grads, exposed in res:
        for k, g in grads.items():
            want = sd_o[k].grad
            err = float((g - want).abs().max()) / (float(want.abs().max()) + 1e-12)
            assert err < 2e-3, (rank, k, err)
    assert torch.equal(res[0][1]['query_feat.weight'], res[1][1]['query_feat.weight'])
