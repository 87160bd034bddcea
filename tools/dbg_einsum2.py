"""Debug helper (GPU box): time cgg_mask_einsum alone under the CGG_TC_DBGMODE experiments."""
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
rt = head._runtime(dev)
torch.cuda.synchronize()
out = torch.empty((10, B, Q, 256, 256), dtype=torch.bfloat16, device=dev)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        rt.mask_einsum(mfd, out)
    e1.record()
    torch.cuda.synchronize()
print('einsum ms', e0.elapsed_time(e1) / 10, 'mode', os.environ.get('CGG_TC_DBGMODE'))
