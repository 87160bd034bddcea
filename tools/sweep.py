"""BASELINE.json configs[4]: masked cross-attention + mask-einsum microbenchmark sweep (GPU box).

queries 100-300, images 512^2-1536^2, key counts per level 256-36,864, attention-mask density 5-100 %.
Attention masks: one contiguous window per query AND i.i.d. per key.
Prints one JSON object per line and writes gpurun_out/sweep.jsonl.  bf16 mode, B=16 (B=8 at 1536^2 x Q=300)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict

dev = torch.device('cuda', 0)
C, HEADS = 256, 8
out_path = os.path.join(ROOT, 'gpurun_out', 'sweep.jsonl')
os.makedirs(os.path.dirname(out_path), exist_ok=True)
fout = open(out_path, 'w')


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    fout.write(line + '\n')


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def packbits(masked):            # (B,Q,K) bool -> (B,Q,ceil(K/32)) int32, bit = 1 masked
    B, Q, K = masked.shape
    W = (K + 31) // 32
    pad = torch.zeros((B, Q, W * 32), dtype=torch.bool, device=masked.device)
    pad[..., :K] = masked
    w = (pad.view(B, Q, W, 32).to(torch.int64) << torch.arange(32, device=masked.device)).sum(-1)
    return (w & 0xffffffff).to(torch.int64).where(w < 2 ** 31, w - 2 ** 32).to(torch.int32).contiguous()


for Q in (100, 200, 300):
    sd = synth.make_params(seed=0, num_queries=Q)
    for size in (512, 1024, 1536):
        B = 8 if (size == 1536 and Q == 300) else 16
        mf, mems = synth.make_inputs(0, B, size, size, dtype=torch.bfloat16)
        head = build_head_from_state_dict(sd, Q, 49, 'bf16', dev, cuda_graph=True)
        mfd, memd = mf.to(dev), [m.to(dev) for m in mems]
        ms_fwd = timeit(lambda: head.decoder_forward(mfd, memd), reps=5)
        emit(bench='decoder_forward', Q=Q, size=size, batch=B, ms=ms_fwd, images_per_s=B / ms_fwd * 1e3)
        rt = head._runtime(dev)
        H4 = size // 4
        mask_out = torch.empty((10, B, Q, H4, H4), dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: rt.mask_einsum(mfd, mask_out))
        flops = 2.0 * Q * C * H4 * H4 * B * 10
        emit(bench='mask_einsum', Q=Q, size=size, batch=B, ms=ms, tflops=flops / ms / 1e9,
             hbm_write_gb_per_s=mask_out.numel() * 2 / ms / 1e6)
        del mask_out
        # masked cross-attention at this size's three key counts, densities 5 % .. 100 % (blob-shaped: one window per query)
        g = torch.Generator(device='cpu').manual_seed(Q + size)
        for lvl, ratio in enumerate((32, 16, 8)):
            K = (size // ratio) ** 2
            q = (torch.randn((B, Q, C), generator=g) * 0.4).to(dev)
            k = torch.randn((B, K, C), generator=g).bfloat16().to(dev)
            v = torch.randn((B, K, C), generator=g).bfloat16().to(dev)
            for kind in ('window', 'iid'):
                for density in (0.05, 0.5, 0.9, 1.0):
                    if kind == 'iid' and density == 1.0:
                        continue
                    # density = fraction of keys MASKED for a query.  'window': the unmasked keys are one contiguous window
                    # per query (blob-shaped, what a mask prediction looks like); 'iid': every key masked independently
                    # (no dead tiles at all -- the worst case for tile skipping)
                    if kind == 'window':
                        width = max(1, int(round(K * (1.0 - density))))
                        masked = torch.ones((B, Q, K), dtype=torch.bool, device=dev)
                        if density < 1.0:
                            start = torch.randint(0, K - width + 1, (B, Q), generator=g).to(dev)
                            idx = torch.arange(K, device=dev)[None, None]
                            masked = ~((idx >= start[..., None]) & (idx < start[..., None] + width))
                    else:
                        masked = (torch.rand((B, Q, K), generator=g) < density).to(dev)
                    am = masked.all(-1).to(torch.uint8).contiguous()
                    bits = packbits(masked)
                    ms = timeit(lambda: rt.masked_attention(q, k, v, bits, am), reps=10)
                    eff = float((~masked).float().mean()) if density < 1.0 else 1.0     # fallback rows attend everywhere
                    emit(bench='masked_attention', Q=Q, size=size, batch=B, level=lvl, keys=K, mask_kind=kind,
                         masked_fraction=density, ms=ms, dense_tflops=4.0 * B * HEADS * Q * K * 32 / ms / 1e9,
                         attended_fraction=eff)
        del head, rt, mfd, memd
        torch.cuda.empty_cache()
fout.close()
