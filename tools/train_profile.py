"""Where the training step's time goes (bench.py's train leg, BASELINE configs[3]): host enqueue time against device time,
the per-kernel device time of one step (torch.profiler / CUPTI), and the same step written with plain torch CUDA ops (the
oracle moved to the GPU -- what the reference executes) in fp32, TF32 and bf16 autocast.
  python tools/train_profile.py [B] [fp32|tf32]"""
import sys, time, json
import torch
sys.path.insert(0, '.')
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
from cgg_b200.grounding import grounding_loss, similarity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
TP = sys.argv[2] if len(sys.argv) > 2 else 'tf32'
Q, ncls1, H, W = 200, 118, 1024, 1024
dev = torch.device('cuda', 0)
sd = synth.make_params(seed=0, num_queries=Q, num_classes_p1=ncls1)
head = build_head_from_state_dict(sd, Q, ncls1, 'fp32', dev, train_precision=TP).train()
mf, mems = synth.make_inputs(100, B, H, W)
mf, mems = mf.to(dev), [m.to(dev) for m in mems]
ids, cap_mask, table, lw, lb = synth.make_captions(0, B)
cap = head.extract_word_embeddings(table.to(dev), lw.to(dev), lb.to(dev), ids.to(dev))
cap_mask = cap_mask.to(dev)
g = torch.Generator().manual_seed(0)
labels = torch.randint(0, ncls1, (B, Q), generator=g).to(dev)
targets = (torch.rand((B, Q, H // 4, W // 4), generator=g) > 0.5).to(dev).float()


def step():
    for p in head.parameters():
        p.grad = None
    cls, emb, mask = head.decoder_forward_auto(mf, mems)
    loss = 0.0
    for j in range(len(cls)):
        loss = loss + grounding_loss(emb[j], cap, cap_mask, 10.0, 2.0)
        logits = similarity(emb[j].reshape(B * Q, -1), head.class_embs, 0.1)
        loss = loss + torch.nn.functional.cross_entropy(logits, labels.reshape(-1))
        loss = loss + torch.nn.functional.binary_cross_entropy_with_logits(mask[j], targets)
    loss.backward()
    return loss


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    host = (time.perf_counter() - t0) / n * 1e3
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


out = {}
ms, host = timeit(step)
out['ours_' + TP] = dict(ms_per_step=ms, host_enqueue_ms=host)
print(json.dumps(out), flush=True)

from torch.profiler import profile, ProfilerActivity
from cgg_b200 import train as _train
calls = []
_orig_gemm = _train._K.gemm


def _logged_gemm(self, A, sA, W, sW, Cc, sC, M, N, K, batch=1, **kw):
    calls.append((M, N, K, batch, 'm' if sA[1] == 1 and sA[2] != 1 else 'k', 'm' if sW[1] == 1 and sW[2] != 1 else 'k'))
    return _orig_gemm(self, A, sA, W, sW, Cc, sC, M, N, K, batch=batch, **kw)


_train._K.gemm = _logged_gemm
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
_train._K.gemm = _orig_gemm
evs = [e for e in prof.events() if e.device_type.name == 'CUDA' and ('gemm_tf32_kernel' in e.name or 'gemm_f32_kernel' in e.name)]
evs.sort(key=lambda e: e.time_range.start)
if len(evs) != len(calls):      # cgg_similarity launches the FMA kernel on its own: keep the tensor-core launches only
    evs = [e for e in evs if 'gemm_tf32_kernel' in e.name]
if len(evs) == len(calls):
    agg = {}
    for c, e in zip(calls, evs):
        key = c + ('simt' if 'gemm_f32' in e.name else 'tf32',)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += e.device_time_total / 1e3
    print('GEMM calls by shape (M, N, K, batch, A major, W major, kernel): count, total ms, us each')
    for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('  %-48s n=%3d %7.3f ms %7.1f us' % (key, n, t, 1e3 * t / n))
else:
    print('gemm events %d != calls %d' % (len(evs), len(calls)))
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print('device time of one step: %.2f ms in %d launches' % (tot, sum(r[2] for r in rows)))
for k, t, c in rows[:25]:
    print('  %-70s n=%4d %8.3f ms %5.1f%%' % (k[:70], c, t, 100 * t / tot))

# ---- the same step with plain torch CUDA ops
from oracle import cgg_oracle as O
sd_o = {k: v.to(dev).requires_grad_(k != 'class_embs') for k, v in sd.items()}


def torch_step():
    for v in sd_o.values():
        v.grad = None
    ref = O.decoder_forward(sd_o, mf, mems)
    loss = 0.0
    for j in range(10):
        loss = loss + O.grounding_loss(ref['emb'][j], cap, cap_mask, 10.0, 2.0)
        logits = O.cls_emb_logits(ref['emb'][j].reshape(B * Q, -1), sd_o['class_embs'], 10.0)
        loss = loss + torch.nn.functional.cross_entropy(logits, labels.reshape(-1))
        loss = loss + torch.nn.functional.binary_cross_entropy_with_logits(ref['mask'][j], targets)
    loss.backward()
    return loss


torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
ms, host = timeit(torch_step, 3)
out['torch_fp32'] = dict(ms_per_step=ms, host_enqueue_ms=host)
torch.backends.cuda.matmul.allow_tf32 = True
ms, host = timeit(torch_step, 3)
out['torch_tf32'] = dict(ms_per_step=ms, host_enqueue_ms=host)


def autocast_step():
    with torch.autocast('cuda', dtype=torch.bfloat16):
        return torch_step()


ms, host = timeit(autocast_step, 3)
out['torch_bf16_autocast'] = dict(ms_per_step=ms, host_enqueue_ms=host)
print(json.dumps(out))
