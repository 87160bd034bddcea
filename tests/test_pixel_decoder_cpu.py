"""CPU tests for row f3 (the pixel decoder before the path): the oracle's restatement of mmdet's MSDeformAttnPixelDecoder /
mmcv's MultiScaleDeformableAttention is pinned on an INDEPENDENT implementation (HuggingFace transformers'
Mask2FormerPixelDecoder carrying the same weights), and the host-side module mirrors mmdet's state_dict layout."""
import pytest
import torch

from cgg_b200 import synth
from oracle import pixel_decoder_oracle as P

CHS = (32, 48, 64, 96)


@pytest.mark.parametrize('size', [(2, 128, 160), (1, 96, 224)])
def test_oracle_matches_the_hf_pixel_decoder(size):
    pytest.importorskip('transformers')
    B, H, W = size
    sd = synth.make_pixel_decoder_params(1, in_channels=CHS)
    feats = synth.make_backbone_feats(1, B, H, W, CHS)
    with torch.no_grad():
        mf, mems = P.pixel_decoder_forward(sd, feats)
        out = P.hf_pixel_decoder(sd, CHS)(feats)
    assert float((out.mask_features - mf).abs().max()) < 2e-5 * float(mf.abs().max())
    for a, b in zip(out.multi_scale_features, mems):
        assert a.shape == b.shape and float((a - b).abs().max()) < 2e-5 * float(b.abs().max())


def test_oracle_is_sensitive_to_its_terms():
    """Dropping the level embedding or the positional term moves the outputs far beyond the pin's tolerance."""
    sd = synth.make_pixel_decoder_params(1, in_channels=CHS)
    feats = synth.make_backbone_feats(1, 1, 96, 128, CHS)
    with torch.no_grad():
        mf, _ = P.pixel_decoder_forward(sd, feats)
        sd2 = dict(sd)
        sd2['level_encoding.weight'] = torch.zeros_like(sd['level_encoding.weight'])
        mf2, _ = P.pixel_decoder_forward(sd2, feats)
    assert float((mf - mf2).abs().max()) > 1e-2 * float(mf.abs().max())


def test_module_mirrors_mmdet_state_dict():
    from cgg_b200.pixel_decoder import MSDeformAttnPixelDecoderB200
    chs = (256, 512, 1024, 2048)
    sd = synth.make_pixel_decoder_params(0, in_channels=chs)
    # the constructor takes the reference's own config block (configs/instance/coco_b48n17.py:38-70)
    m = MSDeformAttnPixelDecoderB200(
        in_channels=list(chs), strides=[4, 8, 16, 32], feat_channels=256, out_channels=256, num_outs=3,
        norm_cfg=dict(type='GN', num_groups=32), act_cfg=dict(type='ReLU'),
        encoder=dict(type='DetrTransformerEncoder', num_layers=6, transformerlayers=dict(
            type='BaseTransformerLayer',
            attn_cfgs=dict(type='MultiScaleDeformableAttention', embed_dims=256, num_heads=8, num_levels=3, num_points=4,
                           im2col_step=64, dropout=0.0, batch_first=False, norm_cfg=None, init_cfg=None),
            ffn_cfgs=dict(type='FFN', embed_dims=256, feedforward_channels=1024, num_fcs=2, ffn_drop=0.0,
                          act_cfg=dict(type='ReLU', inplace=True)),
            operation_order=('self_attn', 'norm', 'ffn', 'norm')), init_cfg=None),
        positional_encoding=dict(type='SinePositionalEncoding', num_feats=128, normalize=True), init_cfg=None)
    own = m.state_dict()
    assert set(own) == set(sd)
    for k in sd:
        assert tuple(own[k].shape) == tuple(sd[k].shape), k
    # MSDeformAttnPixelDecoder of the R50 config: input/lateral/output convs + 6 encoder layers + level embedding + mask conv
    assert sum(v.numel() for v in own.values()) == sum(v.numel() for v in sd.values())
    with pytest.raises(Exception):
        m([torch.zeros(1, c, 8, 8) for c in chs])          # CPU tensors: no fallback
