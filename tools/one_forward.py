"""GPU-box helper for ncu captures: TWO eager B=16 1024^2 bf16 forwards of the whole path (the first one also runs
cgg_prepare); the second forward is launches [N-165, N) of the process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
B, Q = 16, int(os.environ.get('Q', 100))
dev = torch.device('cuda', 0)
sd = synth.make_params(seed=0, num_queries=Q)
head = build_head_from_state_dict(sd, Q, 49, 'bf16', dev)
mf, mems = synth.make_inputs(0, B, 1024, 1024, dtype=torch.bfloat16)
mf, mems = mf.to(dev), [m.to(dev) for m in mems]
for _ in range(2):
    head.decoder_forward(mf, mems)
torch.cuda.synchronize()
