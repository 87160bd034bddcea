"""Second, independent pin of the oracle: HuggingFace `transformers` ships its own port of
Mask2Former's masked-attention decoder (`Mask2FormerMaskedAttentionDecoder`), written by other people from the
original Detectron2 code -- not from mmcv/mmdet and not from this repo.  With the reference head's weights copied in
(state_dict key mapping below), identical inputs must give the oracle's per-layer mask logits, attention masks and
positional encoding.  This guards the one thing the golden fixtures cannot: that `oracle/ref_shim.py`'s restatement of
the un-vendored mmcv 1.7.1 / mmdet 2.28.2 bricks (MultiheadAttention wrapper, BaseTransformerLayer order, FFN,
SinePositionalEncoding) has the right semantics.  CPU only."""
import pytest
import torch

from oracle import cgg_oracle as O
from cgg_b200 import synth

hf = pytest.importorskip('transformers.models.mask2former.modeling_mask2former')
from transformers import Mask2FormerConfig  # noqa: E402


def _hf_decoder(sd, C=256):
    cfg = Mask2FormerConfig(hidden_dim=C, num_attention_heads=8, dim_feedforward=2048, decoder_layers=10, dropout=0.0,
                            pre_norm=False, mask_feature_size=C)
    dec = hf.Mask2FormerMaskedAttentionDecoder(cfg).eval()
    with torch.no_grad():
        dec.layernorm.weight.copy_(sd['transformer_decoder.post_norm.weight'])
        dec.layernorm.bias.copy_(sd['transformer_decoder.post_norm.bias'])
        mlp = dec.mask_predictor.mask_embedder
        lin = [m for m in mlp.modules() if isinstance(m, torch.nn.Linear)]
        assert len(lin) == 3
        for n, i in enumerate((0, 2, 4)):
            lin[n].weight.copy_(sd['mask_embed.%d.weight' % i])
            lin[n].bias.copy_(sd['mask_embed.%d.bias' % i])
        for i, layer in enumerate(dec.layers):
            p = 'transformer_decoder.layers.%d.' % i
            a0, a1 = p + 'attentions.0.attn.', p + 'attentions.1.attn.'
            layer.cross_attn.in_proj_weight.copy_(sd[a0 + 'in_proj_weight'])
            layer.cross_attn.in_proj_bias.copy_(sd[a0 + 'in_proj_bias'])
            layer.cross_attn.out_proj.weight.copy_(sd[a0 + 'out_proj.weight'])
            layer.cross_attn.out_proj.bias.copy_(sd[a0 + 'out_proj.bias'])
            W, b = sd[a1 + 'in_proj_weight'], sd[a1 + 'in_proj_bias']
            for n, proj in enumerate((layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj)):
                proj.weight.copy_(W[n * C:(n + 1) * C])
                proj.bias.copy_(b[n * C:(n + 1) * C])
            layer.self_attn.out_proj.weight.copy_(sd[a1 + 'out_proj.weight'])
            layer.self_attn.out_proj.bias.copy_(sd[a1 + 'out_proj.bias'])
            for n, ln in enumerate((layer.cross_attn_layer_norm, layer.self_attn_layer_norm, layer.final_layer_norm)):
                ln.weight.copy_(sd[p + 'norms.%d.weight' % n])
                ln.bias.copy_(sd[p + 'norms.%d.bias' % n])
            layer.fc1.weight.copy_(sd[p + 'ffns.0.layers.0.0.weight'])
            layer.fc1.bias.copy_(sd[p + 'ffns.0.layers.0.0.bias'])
            layer.fc2.weight.copy_(sd[p + 'ffns.0.layers.1.weight'])
            layer.fc2.bias.copy_(sd[p + 'ffns.0.layers.1.bias'])
    return dec


def test_positional_encoding_matches_hf():
    pe = hf.Mask2FormerSinePositionEmbedding(num_pos_feats=128, normalize=True)
    for (h, w) in [(8, 10), (33, 25)]:
        want = pe(torch.Size((1, 256, h, w)), 'cpu', torch.float32)[0].flatten(1).t()     # (h*w, 256)
        got = O.sine_pos_enc(h, w, 128)
        assert float((got - want).abs().max()) < 1e-6


@pytest.mark.parametrize('B,H,W,pseed,iseed', [(2, 256, 320, 17, 3), (1, 352, 288, 18, 4)])
def test_oracle_matches_hf_masked_attention_decoder(B, H, W, pseed, iseed):
    Q, C = 24, 256
    sd = synth.make_params(seed=pseed, num_queries=Q, perturb=True)
    mf, mems = synth.make_inputs(iseed, B, H, W)
    with torch.no_grad():
        ref = O.decoder_forward(sd, mf, mems)
        dec = _hf_decoder(sd)
        pe = hf.Mask2FormerSinePositionEmbedding(num_pos_feats=C // 2, normalize=True)
        enc, pos, sizes = [], [], []
        for l, m in enumerate(mems):
            enc.append(m.flatten(2).permute(2, 0, 1) + sd['level_embed.weight'][l])           # (K,B,C), head.py:792-796
            pos.append(pe(m.shape, 'cpu', torch.float32).flatten(2).permute(2, 0, 1))          # head.py:798-804
            sizes.append(tuple(m.shape[-2:]))
        out = dec(inputs_embeds=sd['query_feat.weight'][:, None].repeat(1, B, 1),
                  multi_stage_positional_embeddings=pos, pixel_embeddings=mf, encoder_hidden_states=enc,
                  query_position_embeddings=sd['query_embed.weight'][:, None].repeat(1, B, 1),
                  feature_size_list=sizes, return_dict=True)
    assert len(out.masks_queries_logits) == 10
    for j in range(10):
        d = float((out.masks_queries_logits[j] - ref['mask'][j]).abs().max())
        assert d < 2e-4, (j, d)                 # two fp32 CPU implementations, different op order only
        # the decoder state fed to each head call (HF keeps the post_norm'ed copy)
        z = O.layer_norm(ref['x'][j], sd['transformer_decoder.post_norm.weight'], sd['transformer_decoder.post_norm.bias'])
        assert float((out.intermediate_hidden_states[j].transpose(0, 1) - z).abs().max()) < 2e-4


def test_hf_crosscheck_is_sensitive():
    """Dropping the key positional encoding on one side must show up (the check is not vacuous)."""
    Q, B = 24, 1
    sd = synth.make_params(seed=17, num_queries=Q, perturb=True)
    mf, mems = synth.make_inputs(3, B, 256, 320)
    with torch.no_grad():
        ref = O.decoder_forward(sd, mf, mems)
        dec = _hf_decoder(sd)
        enc = [m.flatten(2).permute(2, 0, 1) + sd['level_embed.weight'][l] for l, m in enumerate(mems)]
        pos = [torch.zeros_like(e) for e in enc]
        out = dec(inputs_embeds=sd['query_feat.weight'][:, None].repeat(1, B, 1),
                  multi_stage_positional_embeddings=pos, pixel_embeddings=mf, encoder_hidden_states=enc,
                  query_position_embeddings=sd['query_embed.weight'][:, None].repeat(1, B, 1),
                  feature_size_list=[tuple(m.shape[-2:]) for m in mems], return_dict=True)
    assert float((out.masks_queries_logits[9] - ref['mask'][9]).abs().max()) > 1e-2
