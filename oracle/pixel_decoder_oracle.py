"""TEST INFRASTRUCTURE ONLY (the oracle for SURVEY.md section 8 row f3) -- never imported by the product path.

CPU fp32 restatement, in plain torch ops, of the step BEFORE the decoder-head path:
`mask_features, multi_scale_memorys = self.pixel_decoder(feats)` (open_set/models/mask2former_head.py:787), built from
`pixel_decoder=dict(type='MSDeformAttnPixelDecoder', ...)` (configs/instance/coco_b48n17.py:38-70; same block in every
shipped config).  The class itself is THIRD-PARTY code that is absent from /root/reference: mmdet 2.28.2
`mmdet/models/plugins/msdeformattn_pixel_decoder.py` (MSDeformAttnPixelDecoder.forward) over mmcv-full 1.7.1
`mmcv/ops/multi_scale_deform_attn.py` (MultiScaleDeformableAttention.forward and its pure-torch twin
`multi_scale_deformable_attn_pytorch`), `mmcv/cnn/bricks/conv_module.py` (ConvModule: conv -> GN -> act) and
`mmcv/cnn/bricks/transformer.py` (BaseTransformerLayer with operation_order ('self_attn','norm','ffn','norm'), FFN with
identity add).  Their published algorithm is restated here, function by function.

Parity pinning: neither mmdet nor mmcv can be imported in the build container, and the reference holds no test or golden
vector for this step, so this restatement is pinned on an INDEPENDENT implementation of the same published algorithm:
HuggingFace `transformers` `Mask2FormerPixelDecoder` (a port of the original detectron2 Mask2Former pixel decoder that
mmdet's class also ports) with this file's weights copied in (`tests/test_pixel_decoder_cpu.py`, `hf_pixel_decoder`
below).  With respect to the reference's own third-party dependency that is "parity unpinned" in the strict sense of the
task statement (no output of mmdet itself is available); DESIGN.md says so too.

State-dict keys are mmdet's (`pixel_decoder.` prefix stripped):
  input_convs.{i}.conv.{weight,bias}, input_convs.{i}.gn.{weight,bias}            i = 0..2, from the LOWEST resolution up
  encoder.layers.{l}.attentions.0.{sampling_offsets,attention_weights,value_proj,output_proj}.{weight,bias}
  encoder.layers.{l}.ffns.0.layers.0.0.{weight,bias}, encoder.layers.{l}.ffns.0.layers.1.{weight,bias}
  encoder.layers.{l}.norms.{0,1}.{weight,bias}
  level_encoding.weight (3, C)
  lateral_convs.{i}.conv.weight, lateral_convs.{i}.gn.{weight,bias}, output_convs.{i}.conv.weight, output_convs.{i}.gn.*
  mask_feature.{weight,bias}
"""
import torch
import torch.nn.functional as F

from . import cgg_oracle as O

HEADS, LEVELS, POINTS = 8, 3, 4


def conv_gn(x, w, b, gw, gb, relu=False, groups=32, padding=0):
    """mmcv ConvModule(norm_cfg=GN32, act_cfg=None|ReLU): conv -> GroupNorm(32, C, eps 1e-5) -> activation."""
    y = F.group_norm(F.conv2d(x, w, b, padding=padding), groups, gw, gb, 1e-5)
    return F.relu(y) if relu else y


def ms_deform_attn_core(value, shapes, sampling_locations, attention_weights):
    """mmcv `multi_scale_deformable_attn_pytorch`: value (B, S, H, D), shapes [(h, w)] per level, sampling_locations
    (B, Nq, H, L, P, 2) as (x, y) in [0, 1], attention_weights (B, Nq, H, L, P) -> (B, Nq, H*D).  Bilinear taps with zero
    padding, align_corners=False (grid = 2 loc - 1)."""
    B, _, H, D = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    value_list = value.split([h * w for h, w in shapes], dim=1)
    grids = 2 * sampling_locations - 1
    sampled = []
    for lvl, (h, w) in enumerate(shapes):
        v = value_list[lvl].flatten(2).transpose(1, 2).reshape(B * H, D, h, w)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(v, g, mode='bilinear', padding_mode='zeros', align_corners=False))
    aw = attention_weights.transpose(1, 2).reshape(B * H, 1, Nq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(B, H * D, Nq)
    return out.transpose(1, 2).contiguous()


def reference_points(shapes):
    """mmdet MlvlPointGenerator(strides).single_level_grid_priors(offset 0.5) / (w * stride, h * stride):
    ((x + 0.5) / w, (y + 0.5) / h) per token, levels concatenated; valid ratios are all one (no padding mask)."""
    pts = []
    for (h, w) in shapes:
        ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing='ij')
        pts.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    return torch.cat(pts, 0)


def ms_deform_attn(sd, pre, x, pos, ref, shapes):
    """mmcv MultiScaleDeformableAttention.forward (batch-first here): query = x + pos, value = x (no pos), identity = x."""
    B, S, C = x.shape
    q = x + pos
    value = O.linear(x, sd[pre + 'value_proj.weight'], sd[pre + 'value_proj.bias']).view(B, S, HEADS, C // HEADS)
    off = O.linear(q, sd[pre + 'sampling_offsets.weight'], sd[pre + 'sampling_offsets.bias']).view(B, S, HEADS, LEVELS, POINTS, 2)
    aw = O.linear(q, sd[pre + 'attention_weights.weight'], sd[pre + 'attention_weights.bias']).view(B, S, HEADS, LEVELS * POINTS)
    aw = aw.softmax(-1).view(B, S, HEADS, LEVELS, POINTS)
    norm = torch.tensor([[w, h] for (h, w) in shapes], dtype=torch.float32, device=x.device)   # offset_normalizer: (w, h) per level
    loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    out = ms_deform_attn_core(value, shapes, loc, aw)
    return O.linear(out, sd[pre + 'output_proj.weight'], sd[pre + 'output_proj.bias']) + x


def encoder_layer(sd, l, x, pos, ref, shapes):
    """BaseTransformerLayer, operation_order ('self_attn', 'norm', 'ffn', 'norm'); FFN = Linear-ReLU-Linear + identity."""
    pre = 'encoder.layers.%d.' % l
    x = ms_deform_attn(sd, pre + 'attentions.0.', x, pos, ref, shapes)
    x = O.layer_norm(x, sd[pre + 'norms.0.weight'], sd[pre + 'norms.0.bias'])
    f = torch.relu(O.linear(x, sd[pre + 'ffns.0.layers.0.0.weight'], sd[pre + 'ffns.0.layers.0.0.bias']))
    x = x + O.linear(f, sd[pre + 'ffns.0.layers.1.weight'], sd[pre + 'ffns.0.layers.1.bias'])
    return O.layer_norm(x, sd[pre + 'norms.1.weight'], sd[pre + 'norms.1.bias'])


def pixel_decoder_forward(sd, feats, num_layers=6, return_debug=False):
    """MSDeformAttnPixelDecoder.forward(feats): feats = backbone maps, HIGHEST resolution first (strides 4, 8, 16, 32).
    Returns (mask_feature (B, C_out, H/4, W/4), [memory 1/32, 1/16, 1/8] each (B, C, h, w))."""
    n_in = len(feats)
    B = feats[0].shape[0]
    tokens, pos, shapes = [], [], []
    for i in range(LEVELS):
        f = feats[n_in - 1 - i]
        pre = 'input_convs.%d.' % i
        p = conv_gn(f, sd[pre + 'conv.weight'], sd[pre + 'conv.bias'], sd[pre + 'gn.weight'], sd[pre + 'gn.bias'])
        h, w = f.shape[-2:]
        shapes.append((h, w))
        tokens.append(p.flatten(2).transpose(1, 2))                                   # (B, hw, C)
        pos.append(O.sine_pos_enc(h, w).to(f.device) + sd['level_encoding.weight'][i][None])     # level_embed + pos_embed
    x = torch.cat(tokens, 1)
    pos = torch.cat(pos, 0)[None]
    ref = reference_points(shapes).to(x.device)
    dbg = {'tokens_in': x}
    for l in range(num_layers):
        x = encoder_layer(sd, l, x, pos, ref, shapes)
    dbg['tokens_out'] = x
    outs = [t.transpose(1, 2).reshape(B, -1, h, w) for t, (h, w) in zip(x.split([h * w for h, w in shapes], 1), shapes)]
    for i in range(n_in - LEVELS - 1, -1, -1):
        lat = conv_gn(feats[i], sd['lateral_convs.%d.conv.weight' % i], None, sd['lateral_convs.%d.gn.weight' % i],
                      sd['lateral_convs.%d.gn.bias' % i])
        y = lat + F.interpolate(outs[-1], size=lat.shape[-2:], mode='bilinear', align_corners=False)
        outs.append(conv_gn(y, sd['output_convs.%d.conv.weight' % i], None, sd['output_convs.%d.gn.weight' % i],
                            sd['output_convs.%d.gn.bias' % i], relu=True, padding=1))
    mask_feature = F.conv2d(outs[-1], sd['mask_feature.weight'], sd['mask_feature.bias'])
    if return_debug:
        return mask_feature, outs[:LEVELS], dbg
    return mask_feature, outs[:LEVELS]


# ------------------------------------------------------------------------------------- the independent pin
def hf_pixel_decoder(sd, in_channels, feat=256, ffn=1024, num_layers=6):
    """HuggingFace transformers' Mask2FormerPixelDecoder carrying the weights of `sd` (the independent implementation this
    oracle is pinned on).  Key mapping: input_convs.i.{conv,gn} -> input_projections.i.{0,1}; attentions.0.* ->
    self_attn.*; norms.{0,1} -> self_attn_layer_norm / final_layer_norm; ffns.0.layers.0.0 / .1 -> fc1 / fc2;
    level_encoding.weight -> level_embed; lateral_convs.0 / output_convs.0 -> adapter_1 / layer_1; mask_feature ->
    mask_projection."""
    from transformers import Mask2FormerConfig
    from transformers.models.mask2former.modeling_mask2former import Mask2FormerPixelDecoder
    cfg = Mask2FormerConfig(feature_size=feat, mask_feature_size=sd['mask_feature.weight'].shape[0], hidden_dim=feat,
                            encoder_feedforward_dim=ffn, encoder_layers=num_layers, num_attention_heads=HEADS,
                            feature_strides=[4, 8, 16, 32], common_stride=4, dropout=0.0)
    m = Mask2FormerPixelDecoder(cfg, feature_channels=list(in_channels)).eval()
    t = {}
    for i in range(LEVELS):
        for a, b in (('conv', '0'), ('gn', '1')):
            for p in ('weight', 'bias'):
                t['input_projections.%d.%s.%s' % (i, b, p)] = sd['input_convs.%d.%s.%s' % (i, a, p)]
    for l in range(num_layers):
        s, d = 'encoder.layers.%d.' % l, 'encoder.layers.%d.' % l
        for n in ('sampling_offsets', 'attention_weights', 'value_proj', 'output_proj'):
            for p in ('weight', 'bias'):
                t[d + 'self_attn.%s.%s' % (n, p)] = sd[s + 'attentions.0.%s.%s' % (n, p)]
        for p in ('weight', 'bias'):
            t[d + 'self_attn_layer_norm.' + p] = sd[s + 'norms.0.' + p]
            t[d + 'final_layer_norm.' + p] = sd[s + 'norms.1.' + p]
            t[d + 'fc1.' + p] = sd[s + 'ffns.0.layers.0.0.' + p]
            t[d + 'fc2.' + p] = sd[s + 'ffns.0.layers.1.' + p]
    t['level_embed'] = sd['level_encoding.weight']
    t['adapter_1.0.weight'] = sd['lateral_convs.0.conv.weight']
    t['adapter_1.1.weight'], t['adapter_1.1.bias'] = sd['lateral_convs.0.gn.weight'], sd['lateral_convs.0.gn.bias']
    t['layer_1.0.weight'] = sd['output_convs.0.conv.weight']
    t['layer_1.1.weight'], t['layer_1.1.bias'] = sd['output_convs.0.gn.weight'], sd['output_convs.0.gn.bias']
    t['mask_projection.weight'], t['mask_projection.bias'] = sd['mask_feature.weight'], sd['mask_feature.bias']
    missing, unexpected = m.load_state_dict(t, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return m
