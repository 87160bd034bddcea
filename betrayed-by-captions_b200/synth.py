"""Seeded synthetic weights and inputs for the decoder-head hot path.

Weights follow the reference's initialisation *distributions* (torch defaults for
nn.Linear / nn.Embedding / nn.LayerNorm, then xavier_normal_ on every >=2-D decoder
parameter -- open_set/models/mask2former_head.py:231-240) but are drawn from an explicit
CPU ``torch.Generator`` so that the build container, the GPU box, the oracle and the golden
fixtures all see bit-identical values for a given seed.  Key names are the reference
head's state_dict keys.
"""
import math

import torch


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _linear(g, sd, name, out_f, in_f):
    b = 1.0 / math.sqrt(in_f)
    sd[name + '.weight'] = _uniform(g, (out_f, in_f), b)
    sd[name + '.bias'] = _uniform(g, (out_f,), b)


def _xavier_normal(g, shape):
    fan_out, fan_in = shape
    std = math.sqrt(2.0 / (fan_in + fan_out))
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def make_params(seed=0, num_queries=100, num_classes_p1=49, num_layers=9, embed=256,
                ffn=2048, d_l=768, perturb=False, mask_bias=0.0):
    """Returns an ordered dict name -> fp32 CPU tensor.

    perturb:   also randomise LayerNorm affine params and the attention biases (they are
               1/0 at init, which would hide bugs in how they are applied).
    mask_bias: added to mask_embed.4.bias direction so that mask logits shift; a NEGATIVE
               value makes masks denser (more keys masked), see ``biased_mask_params``.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    sd['query_embed.weight'] = torch.randn((num_queries, embed), generator=g)
    sd['query_feat.weight'] = torch.randn((num_queries, embed), generator=g)
    sd['level_embed.weight'] = torch.randn((3, embed), generator=g)
    _linear(g, sd, 'cls_embed', num_classes_p1, embed)
    for i in (0, 2, 4):
        _linear(g, sd, 'mask_embed.%d' % i, embed, embed)
    _linear(g, sd, 'v2l_transform', d_l, embed)
    ce = torch.randn((num_classes_p1, d_l), generator=g) * (23.7 / math.sqrt(d_l))
    ce[-1] = 0
    sd['class_embs'] = ce
    sd['transformer_decoder.post_norm.weight'] = torch.ones(embed)
    sd['transformer_decoder.post_norm.bias'] = torch.zeros(embed)
    for i in range(num_layers):
        p = 'transformer_decoder.layers.%d.' % i
        for a in (0, 1):
            q = p + 'attentions.%d.attn.' % a
            sd[q + 'in_proj_weight'] = _xavier_normal(g, (3 * embed, embed))
            sd[q + 'in_proj_bias'] = torch.zeros(3 * embed)
            sd[q + 'out_proj.weight'] = _xavier_normal(g, (embed, embed))
            sd[q + 'out_proj.bias'] = torch.zeros(embed)
        sd[p + 'ffns.0.layers.0.0.weight'] = _xavier_normal(g, (ffn, embed))
        sd[p + 'ffns.0.layers.0.0.bias'] = _uniform(g, (ffn,), 1.0 / math.sqrt(embed))
        sd[p + 'ffns.0.layers.1.weight'] = _xavier_normal(g, (embed, ffn))
        sd[p + 'ffns.0.layers.1.bias'] = _uniform(g, (embed,), 1.0 / math.sqrt(ffn))
        for n in (0, 1, 2):
            sd[p + 'norms.%d.weight' % n] = torch.ones(embed)
            sd[p + 'norms.%d.bias' % n] = torch.zeros(embed)
    if perturb:
        for k in list(sd):
            if 'norm' in k and k.endswith('.weight'):
                sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
            elif ('norm' in k and k.endswith('.bias')) or k.endswith('in_proj_bias') \
                    or k.endswith('out_proj.bias'):
                sd[k] = sd[k] + 0.05 * torch.randn(sd[k].shape, generator=g)
    if mask_bias != 0.0:
        sd['mask_embed.4.bias'] = sd['mask_embed.4.bias'] + mask_bias
    return sd


def level_sizes(height, width):
    """(H/4,W/4) of mask_features and the three memory sizes, low -> high resolution
    (1/32, 1/16, 1/8), for an input padded to a multiple of 32."""
    assert height % 32 == 0 and width % 32 == 0
    return (height // 4, width // 4), [(height // s, width // s) for s in (32, 16, 8)]


def make_inputs(seed, batch, height, width, embed=256, dtype=torch.float32):
    """Synthetic stand-in for the pixel decoder's outputs:
    mask_features ~ N(0,1) (B,C,H/4,W/4) and memories ~ N(0,1) at 1/32, 1/16, 1/8."""
    g = torch.Generator().manual_seed(1000 + seed)
    hw4, lv = level_sizes(height, width)
    mf = torch.randn((batch, embed) + hw4, generator=g).to(dtype)
    mems = [torch.randn((batch, embed) + s, generator=g).to(dtype) for s in lv]
    return mf, mems


def make_captions(seed, batch, max_tokens=35, vocab=30522, d_l=768, force_empty=True):
    """Synthetic noun-id captions (open_set/datasets/coco_open.py:117,326-357 shapes):
    ids in [1000, vocab), length ~ U{0..10}, zero padded to max_tokens; plus a seeded
    N(0, 0.02) BERT-like table with LayerNorm(gamma=1, beta=0) (real weights need network)."""
    g = torch.Generator().manual_seed(2000 + seed)
    ids = torch.zeros((batch, max_tokens), dtype=torch.long)
    mask = torch.zeros((batch, max_tokens), dtype=torch.long)
    for b in range(batch):
        n = int(torch.randint(0, 11, (1,), generator=g))
        if force_empty and b == batch - 1:
            n = 0
        ids[b, :n] = torch.randint(1000, vocab, (n,), generator=g)
        mask[b, :n] = 1
    table = torch.randn((vocab, d_l), generator=g) * 0.02
    return ids, mask, table, torch.ones(d_l), torch.zeros(d_l)


CAPTION_CFG_SMALL = dict(nb_layers=2, input_dim=768, hidden_dim=768, ff_dim=512, nb_heads=8, drop_val=0.0, pre_norm=False,
                         seq_length=35, nb_tokens=1500)


def make_caption_params(seed=0, cfg=None, vocab=None, scale=2.0, eos_bias=2.0):
    """Seeded weights of the caption generator under the reference's state_dict keys (`caption_generator.*`,
    open_set/models/transformers/*.py) plus a small BERT embedding table (`bert_embeddings.*`).  `scale` widens the
    projections beyond torch's default init so that the attention and the beam search have something to decide; `eos_bias`
    lifts the [SEP] logit so that beams of different lengths finish."""
    cfg = dict(cfg or CAPTION_CFG_SMALL)
    g = torch.Generator().manual_seed(seed)
    sd = {}
    C, F_, V = cfg['hidden_dim'], cfg['ff_dim'], vocab or cfg['nb_tokens']
    pre = 'caption_generator.'
    if cfg['input_dim'] != C:
        _linear(g, sd, pre + 'adapter', C, cfg['input_dim'])
    for i in range(cfg['nb_layers']):
        p = pre + 'transformer_decoder.decoders.%d.' % i
        _linear(g, sd, p + 'mha_layer.qkv_layer', 3 * C, C)
        _linear(g, sd, p + 'mha_layer.out_layer', C, C)
        for n in ('to_qry', 'to_key', 'to_val', 'to_out'):
            _linear(g, sd, p + 'crx_layer.' + n, C, C)
        _linear(g, sd, p + 'ffn_layer.linears.0.0', F_, C)
        _linear(g, sd, p + 'ffn_layer.linears.1.0', C, F_)
        for n in ('mha', 'crx', 'ffn'):
            sd[p + 'layer_normalz.%s.1.weight' % n] = 1.0 + 0.2 * torch.randn((C,), generator=g)
            sd[p + 'layer_normalz.%s.1.bias' % n] = 0.1 * torch.randn((C,), generator=g)
    _linear(g, sd, pre + 'generator', cfg['nb_tokens'], C)
    sd[pre + 'generator.bias'][102] += eos_bias
    for k in list(sd):
        if k.endswith('.weight') and 'layer_normalz' not in k:
            sd[k] = sd[k] * scale
    sd['bert_embeddings.word_embeddings.weight'] = torch.randn((V, 768), generator=g) * 0.05
    sd['bert_embeddings.LayerNorm.weight'] = 1.0 + 0.1 * torch.randn((768,), generator=g)
    sd['bert_embeddings.LayerNorm.bias'] = 0.05 * torch.randn((768,), generator=g)
    return sd


def make_pixel_decoder_params(seed=0, in_channels=(256, 512, 1024, 2048), feat=256, out_channels=256, ffn=1024,
                              num_layers=6, heads=8, levels=3, points=4):
    """Seeded weights of mmdet's MSDeformAttnPixelDecoder (configs/instance/coco_b48n17.py:38-70) under mmdet's state_dict
    keys (the reference head holds them under `pixel_decoder.`).  Every tensor is random -- including the sampling-offset
    and attention-weight projections, which mmcv initialises to a fixed grid / zero -- so that no term of the arithmetic
    is hidden by a zero; the offsets come out a few pixels wide, like a trained model's."""
    g = torch.Generator().manual_seed(4000 + seed)
    sd = {}
    n_in = len(in_channels)
    for i in range(levels):
        cin = in_channels[n_in - 1 - i]
        sd['input_convs.%d.conv.weight' % i] = _xavier_normal(g, (feat, cin)).view(feat, cin, 1, 1).contiguous()
        sd['input_convs.%d.conv.bias' % i] = 0.05 * torch.randn(feat, generator=g)
        sd['input_convs.%d.gn.weight' % i] = 1 + 0.1 * torch.randn(feat, generator=g)
        sd['input_convs.%d.gn.bias' % i] = 0.05 * torch.randn(feat, generator=g)
    for l in range(num_layers):
        p = 'encoder.layers.%d.' % l
        a = p + 'attentions.0.'
        sd[a + 'sampling_offsets.weight'] = 0.02 * torch.randn((heads * levels * points * 2, feat), generator=g)
        sd[a + 'sampling_offsets.bias'] = 1.5 * torch.randn(heads * levels * points * 2, generator=g)
        sd[a + 'attention_weights.weight'] = 0.05 * torch.randn((heads * levels * points, feat), generator=g)
        sd[a + 'attention_weights.bias'] = 0.3 * torch.randn(heads * levels * points, generator=g)
        for n in ('value_proj', 'output_proj'):
            sd[a + n + '.weight'] = _xavier_normal(g, (feat, feat))
            sd[a + n + '.bias'] = 0.05 * torch.randn(feat, generator=g)
        sd[p + 'ffns.0.layers.0.0.weight'] = _xavier_normal(g, (ffn, feat))
        sd[p + 'ffns.0.layers.0.0.bias'] = _uniform(g, (ffn,), 1.0 / math.sqrt(feat))
        sd[p + 'ffns.0.layers.1.weight'] = _xavier_normal(g, (feat, ffn))
        sd[p + 'ffns.0.layers.1.bias'] = _uniform(g, (feat,), 1.0 / math.sqrt(ffn))
        for n in (0, 1):
            sd[p + 'norms.%d.weight' % n] = 1 + 0.1 * torch.randn(feat, generator=g)
            sd[p + 'norms.%d.bias' % n] = 0.05 * torch.randn(feat, generator=g)
    sd['level_encoding.weight'] = torch.randn((levels, feat), generator=g)
    for i in range(n_in - levels):
        sd['lateral_convs.%d.conv.weight' % i] = _xavier_normal(g, (feat, in_channels[i])).view(feat, in_channels[i], 1, 1).contiguous()
        sd['output_convs.%d.conv.weight' % i] = (torch.randn((feat, feat, 3, 3), generator=g) * math.sqrt(2.0 / (9 * feat)))
        for n in ('lateral_convs', 'output_convs'):
            sd['%s.%d.gn.weight' % (n, i)] = 1 + 0.1 * torch.randn(feat, generator=g)
            sd['%s.%d.gn.bias' % (n, i)] = 0.05 * torch.randn(feat, generator=g)
    sd['mask_feature.weight'] = _xavier_normal(g, (out_channels, feat)).view(out_channels, feat, 1, 1).contiguous()
    sd['mask_feature.bias'] = 0.05 * torch.randn(out_channels, generator=g)
    return sd


def make_backbone_feats(seed, batch, height, width, in_channels=(256, 512, 1024, 2048), dtype=torch.float32):
    """Synthetic backbone maps at strides 4, 8, 16, 32 (highest resolution first), post-ReLU-like (non-negative)."""
    g = torch.Generator().manual_seed(5000 + seed)
    return [torch.randn((batch, c, height // s, width // s), generator=g).clamp_(min=0).to(dtype)
            for c, s in zip(in_channels, (4, 8, 16, 32))]
