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
    # OSPS open-set panoptic head (configs/openset_panoptic/coco_panoptic_p20.py: 64 known things + 53 stuff + 1 = 118
    # class rows) with the 200 queries BASELINE.json asks for and the REAL class embeddings of the reference
    # (datasets/embeddings/coco_panoptic_class_with_bert_emb.json minus datasets/unknown/unknown_p20.txt)
    'osps_q200': dict(batch=1, num_queries=200, height=128, width=160, pseed=8, iseed=5, perturb=True, shift=0.0,
                      ncls1=118, class_embs='coco_panoptic_p20'),
}
MASK_SAMPLE_STRIDE = 5
GRAD_SAMPLE_STRIDE = 8

# name -> (B, Q, caption lengths); includes empty captions and a full (35-token) one
GROUNDING_CASES = {'b6': (6, 20, [3, 0, 35, 7, 1, 10]), 'b3_q100': (3, 100, [5, 9, 2]),
                   'b4_two_empty': (4, 12, [0, 4, 0, 6])}


def real_class_embs(name):
    """(ncls+1, 768) class-embedding buffer exactly as the reference head builds it from its json files
    (mask2former_head.py:202-217); committed fixture written by make_class_embs.py."""
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'class_embs.npz'))
    return torch.from_numpy(z[name].astype('float32'))


def case_tensors(c):
    sd = synth.make_params(seed=c['pseed'], num_queries=c['num_queries'], perturb=c['perturb'],
                           num_classes_p1=c.get('ncls1', 49))
    if c.get('class_embs'):
        sd['class_embs'] = real_class_embs(c['class_embs'])
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
