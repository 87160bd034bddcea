"""Times cgg_gemm_f32 (tf32 form) on the shapes of the training step; `ncu` target for the kernel.
  python tools/tf32_shapes.py [case ...]     cases: lin, ein, dme, df, dw, kv"""
import sys
import torch
sys.path.insert(0, '.')
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
from cgg_b200.train import _K

dev = torch.device('cuda', 0)
head = build_head_from_state_dict(synth.make_params(seed=1, num_queries=8), 8, 49, 'fp32', dev)
k = _K(head._runtime(dev), tf32=True)
B, Q, C, HW = 2, 200, 256, 65536
x = torch.randn(400, C, device=dev); W = torch.randn(C, C, device=dev); bias = torch.randn(C, device=dev)
res = torch.randn(400, C, device=dev); y = torch.empty(400, C, device=dev)
F = torch.randn(B, C, HW, device=dev); me = torch.randn(B, Q, C, device=dev)
mask = torch.empty(B, Q, HW, device=dev); dme = torch.empty(B, Q, C, device=dev); dF = torch.empty(B, C, HW, device=dev)
xk = torch.randn(32768, C, device=dev); yk = torch.empty(32768, C, device=dev); dW = torch.empty(C, C, device=dev)
CASES = {
    'lin': lambda: k.gemm(x, (0, C, 1), W, (0, C, 1), y, (0, C, 1), 400, C, C, bias=bias, R=res, sR=(0, C, 1), r_mod=400),
    'linplain': lambda: k.gemm(x, (0, C, 1), W, (0, C, 1), y, (0, C, 1), 400, C, C),
    'ein': lambda: k.gemm(F, (C * HW, 1, HW), me, (Q * C, C, 1), mask, (Q * HW, 1, HW), HW, Q, C, batch=B, a_mmajor=True, c_mmajor=True),
    'dme': lambda: k.gemm(mask, (Q * HW, HW, 1), F, (C * HW, HW, 1), dme, (Q * C, C, 1), Q, C, HW, batch=B),
    'df': lambda: k.gemm(mask, (Q * HW, 1, HW), me, (Q * C, 1, C), dF, (C * HW, 1, HW), HW, C, Q, batch=B, a_mmajor=True, c_mmajor=True),
    'dw': lambda: k.gemm(y, (0, 1, C), x, (0, 1, C), dW, (0, C, 1), C, C, 400, a_mmajor=True),
    'kv': lambda: k.gemm(xk, (0, C, 1), W, (0, C, 1), yk, (0, C, 1), 32768, C, C, bias=bias),
    'kvdx': lambda: k.gemm(yk, (0, C, 1), W, (0, 1, C), xk, (0, C, 1), 32768, C, C),
}
names = sys.argv[1:] or list(CASES)
mask.normal_()
for n in names:
    fn = CASES[n]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print('%-9s %8.1f us per call' % (n, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
