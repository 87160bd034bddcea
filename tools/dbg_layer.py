"""Debug helper (GPU box): one bf16 decoder layer at a small batch with CGG_DEBUG_SYNC=1."""
import os, sys
os.environ['CGG_DEBUG_SYNC'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (tools/ sits next to tests/)
sys.path.insert(0, ROOT)
import torch
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict

B, Q, H, W = int(sys.argv[1]), int(sys.argv[2]), 256, 256
dev = torch.device('cuda', 0)
sd = synth.make_params(seed=31, num_queries=Q, perturb=True)
mf, mems = synth.make_inputs(7, B, H, W)
head = build_head_from_state_dict(sd, Q, 49, 'bf16', dev)
rt = head._runtime(dev)
sizes = [tuple(m.shape[-2:]) for m in mems]
rt.prepare(mf.shape[2], mf.shape[3], sizes, B)
rt.kv_project([m.to(dev).bfloat16() for m in mems])
torch.cuda.synchronize(); print('kv ok', flush=True)
x = torch.randn(B, Q, 256, device=dev)
out = rt.head_call(x, mf.to(dev).bfloat16(), 0)
torch.cuda.synchronize(); print('head ok', flush=True)
for layer in range(3):
    K = sizes[layer][0] * sizes[layer][1]
    bm = torch.zeros((B, Q, (K + 31) // 32), dtype=torch.int32, device=dev)
    am = torch.zeros((B, Q), dtype=torch.uint8, device=dev)
    y = rt.decoder_layer(layer, x, bm, am)
    torch.cuda.synchronize(); print('layer', layer, 'ok', float(y.abs().max()), flush=True)
