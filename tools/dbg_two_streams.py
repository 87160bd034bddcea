"""Debug helper (GPU box): throughput with two batches in flight (two heads, two streams, alternating)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (tools/ sits next to tests/)
sys.path.insert(0, ROOT)
import torch
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
B, Q = 16, 100
dev = torch.device('cuda', 0)
sd = synth.make_params(seed=0, num_queries=Q)
mf, mems = synth.make_inputs(0, B, 1024, 1024, dtype=torch.bfloat16)
N = int(os.environ.get('NSTREAMS', '2'))
heads = [build_head_from_state_dict(sd, Q, 49, 'bf16', dev, cuda_graph=True) for _ in range(N)]
ins = [(mf.to(dev), [m.to(dev) for m in mems]) for _ in range(N)]
streams = [torch.cuda.Stream() for _ in range(N)]
for i in range(N):
    with torch.cuda.stream(streams[i]):
        for _ in range(4):
            heads[i].decoder_forward(*ins[i])
torch.cuda.synchronize()
K = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in streams:
    s.wait_event(e0)
for k in range(K):
    with torch.cuda.stream(streams[k % N]):
        heads[k % N].decoder_forward(*ins[k % N])
for s in streams:
    torch.cuda.current_stream().wait_stream(s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print('streams %d: %.3f ms per step, %.0f images/s' % (N, ms, B / ms * 1e3))
