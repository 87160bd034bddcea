"""Pixel decoder training form (forward + backward, 2 x 1024^2, R50 widths): per-kernel device time (torch.profiler)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from cgg_b200 import synth
from cgg_b200.pixel_decoder import build_pixel_decoder_from_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
chs = (256, 512, 1024, 2048)
dev = torch.device('cuda', 0)
sd = synth.make_pixel_decoder_params(0, in_channels=chs)
feats = [f.to(dev).requires_grad_(True) for f in synth.make_backbone_feats(0, B, 1024, 1024, chs)]
m = build_pixel_decoder_from_state_dict(sd, chs, dev, precision='tf32').train()


def step():
    for p in m.parameters():
        p.grad = None
    mf, mems = m(feats)
    (mf.square().mean() + sum(t.square().mean() for t in mems)).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
print('pixel decoder fwd+bwd B=%d: %.2f ms' % (B, e0.elapsed_time(e1) / 3))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:22]
tot = sum(r.device_time_total for r in prof.key_averages() if r.device_type.name == 'CUDA')
print('device total %.2f ms' % (tot / 1e3))
for r in rows:
    if r.device_type.name == 'CUDA':
        print('  %-72s %5d x  %8.3f ms' % (r.key[:72], r.count, r.device_time_total / 1e3))
