"""Debug helper (GPU box): SM-clock timeline of the attention kernel inside a real forward (build the library with
CGG_NVCC_EXTRA=-DCGG_AT_TRACING, run with CGG_AT_TRACE=<softmax warp id>)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
B, Q = 16, 100
dev = torch.device('cuda', 0)
sd = synth.make_params(seed=0, num_queries=Q)
head = build_head_from_state_dict(sd, Q, 49, 'bf16', dev)
mf, mems = synth.make_inputs(0, B, 1024, 1024, dtype=torch.bfloat16)
head.decoder_forward(mf.to(dev), [m.to(dev) for m in mems])
torch.cuda.synchronize()
