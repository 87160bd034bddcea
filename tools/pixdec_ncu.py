"""ncu target: one warm forward of the pixel decoder at 16 x 1024^2 (R50 widths), then one forward inside the profiler range.
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gemm_tf32|ms_deform' -c 10 \
      -o gpurun_out/pixdec python tools/pixdec_ncu.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgg_b200 import synth
from cgg_b200.pixel_decoder import build_pixel_decoder_from_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chs = (256, 512, 1024, 2048)
dev = torch.device('cuda', 0)
sd = synth.make_pixel_decoder_params(0, in_channels=chs)
feats = [f.to(dev) for f in synth.make_backbone_feats(0, B, 1024, 1024, chs)]
m = build_pixel_decoder_from_state_dict(sd, chs, dev, precision='tf32').eval()
with torch.no_grad():
    m(feats)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(feats)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
