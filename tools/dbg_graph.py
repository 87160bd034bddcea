"""Debug helper (GPU box): eager vs CUDA-graph replay of one bf16 forward (launch-gap check)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (tools/ sits next to tests/)
sys.path.insert(0, ROOT)
import torch
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
B, Q = 16, 100
dev = torch.device('cuda', 0)
sd = synth.make_params(seed=0, num_queries=Q)
mf, mems = synth.make_inputs(0, B, 1024, 1024, dtype=torch.bfloat16)
head = build_head_from_state_dict(sd, Q, 49, 'bf16', dev)
mfd, memd = mf.to(dev), [m.to(dev) for m in mems]
for _ in range(3):
    out = head.decoder_forward(mfd, memd)
torch.cuda.synchronize()
def timeit(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print('eager ms/step', timeit(lambda: head.decoder_forward(mfd, memd)))
t0 = time.perf_counter()
for _ in range(10): head.decoder_forward(mfd, memd)
print('cpu enqueue ms/step', (time.perf_counter() - t0) * 100)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2): head.decoder_forward(mfd, memd)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        out = head.decoder_forward(mfd, memd)
torch.cuda.synchronize()
print('graph ms/step', timeit(lambda: g.replay()))
