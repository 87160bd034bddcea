"""Debug helper (GPU box): per-tile pipeline trace of the tc attention kernel during one bf16 forward."""
import os, sys
os.environ.setdefault('CGG_AT_TRACE', '2')
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
head.decoder_forward(mf.to(dev), [m.to(dev) for m in mems])
torch.cuda.synchronize()
