"""Debug helper (GPU box): one bf16 forward, then the all-calls mask einsum alone three times
(tc_gemm launches #97..#99 of the process) -- target for `ncu -k regex:tc_gemm_kernel --launch-skip 98`."""
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
head = build_head_from_state_dict(sd, Q, 49, 'bf16', dev)
mfd, memd = mf.to(dev), [m.to(dev) for m in mems]
head.decoder_forward(mfd, memd)
torch.cuda.synchronize()
rt = head._runtime(dev)
out = torch.empty((10, B, Q, 256, 256), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    rt.mask_einsum(mfd, out)
torch.cuda.synchronize()
print('done')
