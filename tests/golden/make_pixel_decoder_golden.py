"""Generates tests/golden/pixel_decoder.npz: outputs of an INDEPENDENT implementation of the pixel decoder (HuggingFace
transformers' Mask2FormerPixelDecoder, the port of the original detectron2 Mask2Former pixel decoder that mmdet's
MSDeformAttnPixelDecoder -- the class the reference configures at configs/instance/coco_b48n17.py:38-70 -- also ports) on
seeded inputs, carrying the seeded weights of cgg_b200.synth.make_pixel_decoder_params under mmdet's key names
(oracle/pixel_decoder_oracle.hf_pixel_decoder does the key mapping).  mmdet / mmcv themselves cannot be imported in the
build container, so these vectors pin the oracle -- and through it the CUDA path -- on the published algorithm as a second
party implemented it, not on mmdet's own outputs.

    python tests/golden/make_pixel_decoder_golden.py        (build container; needs `transformers`)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cgg_b200 import synth                                   # noqa: E402
from oracle import pixel_decoder_oracle as P                 # noqa: E402

CASE = dict(seed=3, batch=1, height=64, width=96, in_channels=(16, 24, 32, 48))


def main():
    import transformers
    c = CASE
    sd = synth.make_pixel_decoder_params(c['seed'], in_channels=c['in_channels'])
    feats = synth.make_backbone_feats(c['seed'], c['batch'], c['height'], c['width'], c['in_channels'])
    with torch.no_grad():
        out = P.hf_pixel_decoder(sd, c['in_channels'])(feats)
    arrays = dict(mask_features=out.mask_features.numpy())
    for i, m in enumerate(out.multi_scale_features):
        arrays['memory%d' % i] = m.numpy()
    arrays['generator'] = np.array('transformers %s Mask2FormerPixelDecoder, torch %s' % (transformers.__version__, torch.__version__))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'pixel_decoder.npz')
    np.savez_compressed(path, **arrays)
    print('wrote', path, {k: v.shape for k, v in arrays.items()})


if __name__ == '__main__':
    main()
