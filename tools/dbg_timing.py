"""Debug helper (GPU box): kernel-internal timeline of the tc GEMM launches of one bf16 forward."""
import os, sys
os.environ['CGG_TC_TIMING'] = '1'
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
head.decoder_forward(mfd, memd)
torch.cuda.synchronize()
sys.stderr.write('==== second forward ====\n')
head.decoder_forward(mfd, memd)
torch.cuda.synchronize()
