"""Shared, seeded case definitions for the golden fixtures (used by make_golden.py in the
build container and by the tests everywhere).  Only seeds and shapes live here."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cgg_b200 import synth  # noqa: E402

# name -> case.  `shift`: mask_features channel 0 is set to 1.0 and mask_embed.4.bias[0] is
# shifted, so every mask logit moves by `shift`; negative = denser masks, which makes
# all-masked rows frequent and exercises the fallback (mask2former_head.py:825-826).
HEAD_CASES = {
    'tiny': dict(batch=2, num_queries=16, height=96, width=128, pseed=3, iseed=0, perturb=True, shift=0.0),
    'dense_fallback': dict(batch=2, num_queries=16, height=96, width=128, pseed=4, iseed=1, perturb=True, shift=-0.9),
    'sparse': dict(batch=1, num_queries=16, height=96, width=128, pseed=5, iseed=2, perturb=False, shift=+0.9),
    # ragged: key counts 15/60/240 (not multiples of 32), W/4 = 24, Q = 100 (not a multiple of 8)
    'ragged_q100': dict(batch=1, num_queries=100, height=160, width=96, pseed=6, iseed=3, perturb=True, shift=0.0),
}
MASK_SAMPLE_STRIDE = 5
GRAD_SAMPLE_STRIDE = 8

# name -> (B, Q, caption lengths); includes empty captions and a full (35-token) one
GROUNDING_CASES = {'b6': (6, 20, [3, 0, 35, 7, 1, 10]), 'b3_q100': (3, 100, [5, 9, 2]),
                   'b4_two_empty': (4, 12, [0, 4, 0, 6])}


def case_tensors(c):
    sd = synth.make_params(seed=c['pseed'], num_queries=c['num_queries'], perturb=c['perturb'])
    mf, mems = synth.make_inputs(c['iseed'], c['batch'], c['height'], c['width'])
    if c['shift'] != 0.0:
        mf[:, 0] = 1.0
        sd['mask_embed.4.bias'][0] += c['shift']
    return sd, mf, mems


def param_checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def grounding_tensors(name):
    B, Q, lens = GROUNDING_CASES[name]
    g = torch.Generator().manual_seed(77 + sum(map(ord, name)))
    pred = torch.randn((B, Q, 768), generator=g)
    cap = torch.randn((B, 35, 768), generator=g) * 0.85
    m = torch.zeros((B, 35), dtype=torch.long)
    for b, n in enumerate(lens):
        m[b, :n] = 1
    return pred, cap, m
