"""Debug helper (GPU box): graph-mode forward time vs image size -- the small-image time is the floor set by
the latency-bound linears / LayerNorms / launches of the layer chain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (tools/ sits next to tests/)
sys.path.insert(0, ROOT)
import torch
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
B, Q = 16, 100
dev = torch.device('cuda', 0)
sd = synth.make_params(seed=0, num_queries=Q)
for HW in (256, 512, 1024):
    mf, mems = synth.make_inputs(0, B, HW, HW, dtype=torch.bfloat16)
    head = build_head_from_state_dict(sd, Q, 49, 'bf16', dev, cuda_graph=True)
    mfd, memd = mf.to(dev), [m.to(dev) for m in mems]
    for _ in range(5):
        head.decoder_forward(mfd, memd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        head.decoder_forward(mfd, memd)
    e1.record()
    torch.cuda.synchronize()
    print('size %d: %.3f ms per forward' % (HW, e0.elapsed_time(e1) / 20))
    del head
