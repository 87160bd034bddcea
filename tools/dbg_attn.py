"""Debug helper (GPU box): level-2 cross-attention stage alone (B=16, Q=100, K=16384, 50% random mask)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (tools/ sits next to tests/)
sys.path.insert(0, ROOT)
import torch
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
B, Q, K = 16, 100, 16384
dev = torch.device('cuda', 0)
sd = synth.make_params(seed=0, num_queries=Q)
head = build_head_from_state_dict(sd, Q, 49, 'bf16', dev)
rt = head._runtime(dev)
rt.prepare(64, 64, [(8, 8), (16, 16), (32, 32)], B)
q = torch.randn(B, Q, 256, device=dev) * 0.4
k = torch.randn(B, K, 256, device=dev).bfloat16()
v = torch.randn(B, K, 256, device=dev).bfloat16()
bm = torch.randint(-2**31, 2**31 - 1, (B, Q, K // 32), dtype=torch.int32, device=dev)
am = torch.zeros(B, Q, dtype=torch.uint8, device=dev)
for _ in range(3):
    out = rt.masked_attention(q, k, v, bm, am)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = rt.masked_attention(q, k, v, bm, am)
e1.record(); torch.cuda.synchronize()
print('attention level-2 ms', e0.elapsed_time(e1) / 10)
