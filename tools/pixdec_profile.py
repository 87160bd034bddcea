"""Pixel decoder (row f3) at the BASELINE configs[1] shapes: B images of 1024^2, R50 channel widths.  Whole-forward time
(CUDA events), per-kernel device time (torch.profiler), and the same step as plain torch CUDA ops (the oracle moved to the
GPU: cuDNN convs, ATen GroupNorm / grid_sample -- what mmcv's pure-torch path executes) for comparison.
usage: python tools/pixdec_profile.py [B] [precision] [--no-torch]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgg_b200 import synth
from cgg_b200.pixel_decoder import build_pixel_decoder_from_state_dict
from oracle import pixel_decoder_oracle as P

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else 'tf32'
chs = (256, 512, 1024, 2048)
dev = torch.device('cuda', 0)
sd = synth.make_pixel_decoder_params(0, in_channels=chs)
feats = [f.to(dev) for f in synth.make_backbone_feats(0, B, 1024, 1024, chs)]
m = build_pixel_decoder_from_state_dict(sd, chs, dev, precision=prec).eval()


def timeit(fn, n=5, w=2):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    ms = timeit(lambda: m(feats))
    print('b200 pixel decoder [%s] B=%d: %.3f ms  (%.1f images/s)' % (prec, B, ms, B / ms * 1e3))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(feats)
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:14]
    for r in rows:
        print('  %-70s %5d x  %9.3f ms' % (r.key[:70], r.count, r.device_time_total / 1e3))
    seq = [(e.name, e.device_time_total) for e in prof.events() if e.device_time_total > 0 and 'Memcpy' not in e.name and 'Memset' not in e.name]
    print('  in launch order (ms):', ' '.join('%s:%.2f' % (('G' if 'gemm_tf32' in n else 'F' if 'gemm_f32' in n else 'D' if 'deform' in n else n.split('::')[-1][:6]), t / 1e3) for n, t in seq))
    out = {'B': B, 'precision': prec, 'ms': ms, 'images_per_s': B / ms * 1e3}
    if '--no-torch' not in sys.argv:
        sd_d = {k: v.to(dev) for k, v in sd.items()}
        for name, tf32 in (('fp32', False), ('tf32', True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            t = timeit(lambda: P.pixel_decoder_forward(sd_d, feats), n=3, w=1)
            print('torch CUDA ops [%s]: %.3f ms' % (name, t))
            out['torch_%s_ms' % name] = t
        with torch.autocast('cuda', dtype=torch.bfloat16):
            t = timeit(lambda: P.pixel_decoder_forward(sd_d, feats), n=3, w=1)
        print('torch CUDA ops [bf16 autocast]: %.3f ms' % t)
        out['torch_bf16_autocast_ms'] = t
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/pixdec_profile.json', 'w'))
