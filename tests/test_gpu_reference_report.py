"""Report-style GPU test: times the reference-equivalent PyTorch path (the oracle's plain torch ops,
i.e. what the reference head executes through torch.nn / cuBLAS) ON THE GPU, fp32 and under bf16
autocast, at the benchmark configuration, and writes gpurun_out/torch_gpu_baseline.json.  This is
the "reference PyTorch-GPU decoder-head throughput" BASELINE.md section 3 asks to measure next to
our kernels.  It asserts nothing about speed."""
import json
import os

import pytest
import torch

from oracle import cgg_oracle as O
from cgg_b200 import synth

pytestmark = pytest.mark.gpu


def _time(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_torch_gpu_reference_throughput():
    B, Q = 16, 100
    dev = torch.device('cuda', 0)
    sd = {k: v.to(dev) for k, v in synth.make_params(seed=0, num_queries=Q).items()}
    mf, mems = synth.make_inputs(0, B, 1024, 1024)
    mf, mems = mf.to(dev), [m.to(dev) for m in mems]
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ms32 = _time(lambda: O.decoder_forward(sd, mf, mems))
        with torch.autocast('cuda', dtype=torch.bfloat16):
            ms16 = _time(lambda: O.decoder_forward(sd, mf, mems))
    out = dict(config='B=16, Q=100, 1024x1024, 9 layers', fp32_ms=ms32, fp32_images_per_s=B / ms32 * 1e3,
               bf16_autocast_ms=ms16, bf16_autocast_images_per_s=B / ms16 * 1e3,
               note='oracle restatement of the reference path executed with torch CUDA ops (cuBLAS/ATen)')
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/torch_gpu_baseline.json', 'w') as f:
        json.dump(out, f, indent=1)
    print(out)
    assert ms32 > 0 and ms16 > 0
