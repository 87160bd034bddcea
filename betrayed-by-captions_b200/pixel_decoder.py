"""The step BEFORE the decoder-head path:
`mask_features, multi_scale_memorys = self.pixel_decoder(feats)` (open_set/models/mask2former_head.py:787), i.e. mmdet's
`MSDeformAttnPixelDecoder` as configured at configs/instance/coco_b48n17.py:38-70 (1x1 input convs + GroupNorm, a 6-layer
multi-scale deformable-attention encoder over the three coarse levels, one FPN top-down step to 1/4 scale, the mask-feature
1x1 conv), built on the stage kernels of the C-ABI library -- forward AND backward, so that the gradients of the head's
training step flow on into the backbone features exactly as the reference's autograd does.

`MSDeformAttnPixelDecoderB200` carries mmdet's state_dict keys (the reference head stores them under `pixel_decoder.`) and
the reference's call contract: `forward(feats) -> (mask_feature (B, C, H/4, W/4), [memory 1/32, 1/16, 1/8])`.

B200 layout: activations are TOKEN-MAJOR fp32 (images, pixels, channels) from the first 1x1 conv on -- the NCHW backbone
maps are read in place as MN-major GEMM operands (TMA builds the tiles from the strided tensor), every linear layer is a
K-major product, a deformable-attention tap is one contiguous 128-byte read per head, the 3x3 output conv is an implicit
GEMM whose nine taps are TMA loads with out-of-bounds zero fill (no im2col, no padded copy), and only the four tensors the
head consumes are written NCHW.  PyTorch is the tape, the allocator and the stream; `precision='tf32'` runs every
contraction on tcgen05 `kind::tf32` (csrc/gemm_tf32.cu), 'fp32' on the FMA kernels (the parity mode).
"""
import ctypes as C

import torch
import torch.nn as nn

from . import lib as _lib
from .train import _K, _p, _Linear, _LayerNorm, _AddRows


# ------------------------------------------------------------------------------------------- autograd nodes
def _gemm_conv(k, x, Hh, Ww, W2, y, cin, cout, batch):
    """y (B, H*W, cout) = conv3x3(x (B, H*W, cin)) with W2 (cout, 9*cin), k = tap * cin + c."""
    k.gemm(x, (Hh * Ww * cin, cin, 1), W2, (0, 9 * cin, 1), y, (Hh * Ww * cout, cout, 1), Ww, cout, 9 * cin,
           batch=batch * Hh, batch_inner=Hh, s2=(Ww * cin, 0, Ww * cout), conv_cin=cin)


class _Conv1x1NCHW(torch.autograd.Function):
    """tokens (B, hw, N) = 1x1 conv of an NCHW map x (B, Cin, h, w), read in place as an MN-major operand; W (N, Cin)."""

    @staticmethod
    def forward(ctx, k, x, W, b):
        B, Cin, h, w = x.shape
        P, N = h * w, W.shape[0]
        y = k.new(B, P, N)
        k.gemm(x, (Cin * P, 1, P), W, (0, Cin, 1), y, (P * N, N, 1), P, N, Cin, batch=B, bias=b, a_mmajor=True)
        ctx.k = k
        ctx.save_for_backward(x, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        k = ctx.k
        x, W = ctx.saved_tensors
        B, Cin, h, w = x.shape
        P, N = h * w, W.shape[0]
        dy = dy.contiguous()
        dx = dW = db = None
        if ctx.needs_input_grad[1]:                   # dx[b, c, p] = sum_n dy[b, p, n] W[n, c]   (written NCHW, p contiguous)
            dx = k.new(B, Cin, h, w)
            k.gemm(dy, (P * N, N, 1), W, (0, 1, Cin), dx, (Cin * P, 1, P), P, Cin, N, batch=B, c_mmajor=True)
        if ctx.needs_input_grad[2]:                   # dW[n, c] = sum_b sum_p dy[b, p, n] x[b, c, p]
            part = k.new(B, N, Cin)
            k.gemm(dy, (P * N, 1, N), x, (Cin * P, P, 1), part, (N * Cin, Cin, 1), N, Cin, P, batch=B, a_mmajor=True)
            dW = k.new(N, Cin)
            k.chk(k.lib.cgg_sum_batch(k.h, _p(part), _p(dW), B, N * Cin, k.s()), 'cgg_sum_batch')
        if ctx.needs_input_grad[3]:
            db = k.new(N)
            k.chk(k.lib.cgg_colsum(k.h, _p(dy), _p(db), B * P, N, 1.0, 0, k.s()), 'cgg_colsum')
        return None, dx, dW, db


class _Conv1x1ToNCHW(torch.autograd.Function):
    """NCHW map (B, N, h, w) = 1x1 conv of tokens x (B, hw, Cin): the mask-feature projection, written pixel-contiguous."""

    @staticmethod
    def forward(ctx, k, x, W, b, h, w):
        B, P, Cin = x.shape
        N = W.shape[0]
        y = k.new(B, N, h, w)
        k.gemm(x, (P * Cin, Cin, 1), W, (0, Cin, 1), y, (N * P, 1, P), P, N, Cin, batch=B, bias=b, c_mmajor=True)
        ctx.k = k
        ctx.save_for_backward(x, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        k = ctx.k
        x, W = ctx.saved_tensors
        B, P, Cin = x.shape
        N = W.shape[0]
        dyt = k.new(B, P, N)                          # the gradient as tokens: the three products are then a linear layer's
        k.chk(k.lib.cgg_nchw_to_tokens(k.h, _p(dy.contiguous()), _p(dyt), P * N, B, P, N, 0, k.s()), 'cgg_nchw_to_tokens')
        dx = dW = db = None
        if ctx.needs_input_grad[1]:                   # dx[(b,p), c] = sum_n dy[(b,p), n] W[n, c]
            dx = k.new(B, P, Cin)
            k.gemm(dyt, (0, N, 1), W, (0, 1, Cin), dx, (0, Cin, 1), B * P, Cin, N)
        if ctx.needs_input_grad[2]:                   # dW[n, c] = sum_(b,p) dy[(b,p), n] x[(b,p), c]
            dW = k.new(N, Cin)
            k.gemm(dyt, (0, 1, N), x, (0, 1, Cin), dW, (0, Cin, 1), N, Cin, B * P, a_mmajor=True)
        if ctx.needs_input_grad[3]:
            db = k.new(N)
            k.chk(k.lib.cgg_colsum(k.h, _p(dyt), _p(db), B * P, N, 1.0, 0, k.s()), 'cgg_colsum')
        return None, dx, dW, db, None, None


class _Conv3x3(torch.autograd.Function):
    """3x3 conv (stride 1, padding 1, no bias) over tokens x (B, H*W, Cin) as an implicit GEMM; W2 (Cout, 9*Cin) with
    k = (3*ky + kx) * Cin + c."""

    @staticmethod
    def forward(ctx, k, x, W2, Hh, Ww):
        B, P, Cin = x.shape
        Cout = W2.shape[0]
        y = k.new(B, P, Cout)
        _gemm_conv(k, x, Hh, Ww, W2, y, Cin, Cout, B)
        ctx.k, ctx.hw = k, (Hh, Ww)
        ctx.save_for_backward(x, W2)
        return y

    @staticmethod
    def backward(ctx, dy):
        k = ctx.k
        x, W2 = ctx.saved_tensors
        Hh, Ww = ctx.hw
        B, P, Cin = x.shape
        Cout = W2.shape[0]
        dy = dy.contiguous()
        dx = dW2 = None
        if ctx.needs_input_grad[1]:      # the adjoint is the same convolution with the taps mirrored and the channel roles swapped
            Wf = W2.view(Cout, 9, Cin).flip(1).permute(2, 1, 0).reshape(Cin, 9 * Cout).contiguous()
            dx = k.new(B, P, Cin)
            _gemm_conv(k, dy, Hh, Ww, Wf, dx, Cout, Cin, B)
        if ctx.needs_input_grad[2]:
            # dW2[o, tap, c] = sum_pixels dy[pix, o] x[pix + tap shift, c]: over zero-bordered copies a tap shift is a flat
            # row offset (the border rows of dy are zero, so wrapped terms vanish) -> nine long-K products (split over K)
            xp = torch.nn.functional.pad(x.view(B, Hh, Ww, Cin), (0, 0, 1, 1, 1, 1)).view(-1, Cin)
            dp = torch.nn.functional.pad(dy.view(B, Hh, Ww, Cout), (0, 0, 1, 1, 1, 1)).view(-1, Cout)
            guard = Ww + 3
            Kp = xp.shape[0] - 2 * guard
            dW2 = k.new(Cout, 9 * Cin)
            for tap in range(9):
                sh = (tap // 3 - 1) * (Ww + 2) + (tap % 3 - 1)
                k.gemm(dp[guard:], (0, 1, Cout), xp[guard + sh:], (0, 1, Cin), dW2[:, tap * Cin:], (0, 9 * Cin, 1),
                       Cout, Cin, Kp, a_mmajor=True)
        return None, dx, dW2, None, None


class _GroupNorm(torch.autograd.Function):
    """GroupNorm(groups) + optional ReLU over tokens x (B, P, C) (mmcv ConvModule's norm + activation)."""

    @staticmethod
    def forward(ctx, k, x, gamma, beta, groups, relu):
        B, P, Cc = x.shape
        y, mr = k.new(B, P, Cc), k.new(B, groups, 2)
        nb = k.lib.cgg_group_norm_scratch_bytes(B, P, Cc, groups)
        scratch = torch.empty(max(nb, 4), dtype=torch.uint8, device=k.dev)
        k.chk(k.lib.cgg_group_norm_tokens(k.h, _p(x), _p(gamma), _p(beta), _p(y), _p(mr), _p(scratch), nb, B, P, Cc, groups,
                                          1e-5, int(relu), k.s()), 'cgg_group_norm_tokens')
        ctx.k, ctx.groups, ctx.relu = k, groups, relu
        ctx.save_for_backward(x, gamma, mr, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        k = ctx.k
        x, gamma, mr, y = ctx.saved_tensors
        B, P, Cc = x.shape
        dy = dy.contiguous()
        g = dy
        if ctx.relu:
            g = k.new(B, P, Cc)
            k.chk(k.lib.cgg_relu_backward(k.h, _p(y), _p(dy), _p(g), B * P * Cc, 1.0, k.s()), 'cgg_relu_backward')
        dx, dg, db = k.new(B, P, Cc), k.new(Cc), k.new(Cc)
        nb = k.lib.cgg_group_norm_scratch_bytes(B, P, Cc, ctx.groups)
        scratch = torch.empty(max(nb, 4), dtype=torch.uint8, device=k.dev)
        k.chk(k.lib.cgg_group_norm_tokens_backward(k.h, _p(x), _p(g), _p(mr), _p(gamma), _p(dx), _p(dg), _p(db), _p(scratch), nb,
                                                   B, P, Cc, ctx.groups, k.s()), 'cgg_group_norm_tokens_backward')
        return None, dx, dg, db, None, None


class _MSDeformCore(torch.autograd.Function):
    """mmcv MultiScaleDeformableAttention core: value (B, S, C), offsets (B, S, heads*L*P*2), weight logits (B, S, heads*L*P)."""

    @staticmethod
    def forward(ctx, k, value, off, logits, shapes, heads, points):
        B, S, Cc = value.shape
        L = len(shapes)
        hs = (C.c_int * L)(*[s[0] for s in shapes])
        ws = (C.c_int * L)(*[s[1] for s in shapes])
        out = k.new(B, S, Cc)
        k.chk(k.lib.cgg_ms_deform_attn(k.h, _p(value), Cc, _p(off), off.shape[-1], _p(logits), logits.shape[-1], _p(out), B, S,
                                       heads, L, points, hs, ws, k.s()), 'cgg_ms_deform_attn')
        ctx.k, ctx.geom = k, (hs, ws, heads, L, points)
        ctx.save_for_backward(value, off, logits)
        return out

    @staticmethod
    def backward(ctx, dout):
        k = ctx.k
        value, off, logits = ctx.saved_tensors
        hs, ws, heads, L, points = ctx.geom
        B, S, Cc = value.shape
        dout = dout.contiguous()
        dv, do, dl = k.new(B, S, Cc), torch.zeros_like(off), k.new(*logits.shape)
        k.chk(k.lib.cgg_ms_deform_attn_backward(k.h, _p(value), _p(off), _p(logits), _p(dout), _p(dv), _p(do), _p(dl), B, S,
                                                heads, L, points, hs, ws, k.s()), 'cgg_ms_deform_attn_backward')
        return None, dv, do, dl, None, None, None


class _UpsampleAdd(torch.autograd.Function):
    """lateral (B, H*W, C) + bilinear upsample of the level slice tokens[:, start:start+h*w] (B, S, C)."""

    @staticmethod
    def forward(ctx, k, lat, tokens, start, hw, HW):
        B, S, Cc = tokens.shape
        out = k.new(*lat.shape)
        src = tokens[:, start:start + hw[0] * hw[1]]
        k.chk(k.lib.cgg_upsample_add_tokens(k.h, _p(lat), C.c_void_p(src.data_ptr()), S * Cc, _p(out), B, HW[0], HW[1], hw[0],
                                            hw[1], Cc, k.s()), 'cgg_upsample_add_tokens')
        ctx.k, ctx.geom = k, (B, S, Cc, start, hw, HW)
        return out

    @staticmethod
    def backward(ctx, dout):
        k = ctx.k
        B, S, Cc, start, hw, HW = ctx.geom
        dout = dout.contiguous()
        dtok = None
        if ctx.needs_input_grad[2]:
            dlevel = k.new(B, hw[0] * hw[1], Cc)
            k.chk(k.lib.cgg_upsample_add_tokens_backward(k.h, _p(dout), _p(dlevel), B, HW[0], HW[1], hw[0], hw[1], Cc, k.s()),
                  'cgg_upsample_add_tokens_backward')
            dtok = torch.zeros((B, S, Cc), dtype=torch.float32, device=k.dev)
            dtok[:, start:start + hw[0] * hw[1]] = dlevel
        return None, dout, dtok, None, None, None


class _LevelToNCHW(torch.autograd.Function):
    """One level of the token buffer (B, S, C) -> (B, C, h, w): the memories the decoder head consumes (head.py:789-806)."""

    @staticmethod
    def forward(ctx, k, tokens, start, hw, bf16):
        B, S, Cc = tokens.shape
        P = hw[0] * hw[1]
        out = torch.empty((B, Cc, hw[0], hw[1]), dtype=torch.bfloat16 if bf16 else torch.float32, device=k.dev)
        src = tokens[:, start:start + P]
        k.chk(k.lib.cgg_tokens_to_nchw(k.h, C.c_void_p(src.data_ptr()), S * Cc, _p(out), int(bf16), B, P, Cc, k.s()),
              'cgg_tokens_to_nchw')
        ctx.k, ctx.geom = k, (B, S, Cc, start, P)
        return out

    @staticmethod
    def backward(ctx, dout):
        k = ctx.k
        B, S, Cc, start, P = ctx.geom
        dout = dout.float().contiguous()
        dtok = torch.zeros((B, S, Cc), dtype=torch.float32, device=k.dev)
        dst = dtok[:, start:start + P]
        k.chk(k.lib.cgg_nchw_to_tokens(k.h, _p(dout), C.c_void_p(dst.data_ptr()), S * Cc, B, P, Cc, 0, k.s()), 'cgg_nchw_to_tokens')
        return None, dtok, None, None, None


# ------------------------------------------------------------------------------------------- modules (mmdet's names)
class _ConvGN(nn.Module):
    """mmcv ConvModule(norm_cfg=GN): children `conv` and `gn`."""

    def __init__(self, cin, cout, ksize, bias, groups=32):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, ksize, padding=ksize // 2, bias=bias)
        self.gn = nn.GroupNorm(groups, cout)


class _MSDA(nn.Module):
    def __init__(self, dim, heads, levels, points):
        super().__init__()
        self.sampling_offsets = nn.Linear(dim, heads * levels * points * 2)
        self.attention_weights = nn.Linear(dim, heads * levels * points)
        self.value_proj = nn.Linear(dim, dim)
        self.output_proj = nn.Linear(dim, dim)


class _FFN(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(dim, hidden), nn.ReLU(inplace=True), nn.Dropout(0.0)),
                                    nn.Linear(hidden, dim), nn.Dropout(0.0))


class _EncoderLayer(nn.Module):
    def __init__(self, dim, hidden, heads, levels, points):
        super().__init__()
        self.attentions = nn.ModuleList([_MSDA(dim, heads, levels, points)])
        self.ffns = nn.ModuleList([_FFN(dim, hidden)])
        self.norms = nn.ModuleList([nn.LayerNorm(dim), nn.LayerNorm(dim)])


class _Encoder(nn.Module):
    def __init__(self, n, *a):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(*a) for _ in range(n)])


class _PDRuntime:
    """A C-ABI handle for the stage kernels (no decoder-head state is used)."""

    def __init__(self, device):
        self.lib = _lib.load()
        self.device = device
        self.handle = C.c_void_p()
        cfg = _lib.Config(1, 256, 8, 1, 1, 1, 0, _lib.FP32, 0)
        with torch.cuda.device(device):
            _lib.check(self.lib.cgg_create(C.byref(self.handle), C.byref(cfg)), None, 'cgg_create')

    def __del__(self):
        try:
            if self.handle:
                self.lib.cgg_destroy(self.handle)
        except Exception:
            pass


class MSDeformAttnPixelDecoderB200(nn.Module):
    """Drop-in for mmdet's MSDeformAttnPixelDecoder as the reference configures it (coco_b48n17.py:38-70): same constructor
    keywords (the nested mmcv config dicts are read for the sizes), same state_dict keys, same forward contract."""

    def __init__(self, in_channels=(256, 512, 1024, 2048), strides=(4, 8, 16, 32), feat_channels=256, out_channels=256,
                 num_outs=3, norm_cfg=None, act_cfg=None, encoder=None, positional_encoding=None, init_cfg=None,
                 precision='tf32', out_dtype=torch.float32, **kwargs):
        super().__init__()
        enc = encoder or {}
        tl = enc.get('transformerlayers', {})
        attn = tl.get('attn_cfgs', {})
        ffn = tl.get('ffn_cfgs', {})
        self.num_layers = enc.get('num_layers', 6)
        self.heads = attn.get('num_heads', 8)
        self.levels = attn.get('num_levels', 3)
        self.points = attn.get('num_points', 4)
        hidden = ffn.get('feedforward_channels', 1024)
        groups = (norm_cfg or {}).get('num_groups', 32)
        if feat_channels != self.heads * 32:
            raise ValueError('feat_channels must be heads * 32 (the deformable-attention kernel works on 32-channel heads)')
        if precision not in ('fp32', 'tf32'):
            raise ValueError("precision must be 'fp32' or 'tf32'")
        self.in_channels, self.strides = list(in_channels), list(strides)
        self.feat_channels, self.out_channels, self.num_outs, self.groups = feat_channels, out_channels, num_outs, groups
        self.precision, self.out_dtype = precision, out_dtype
        n_in = len(self.in_channels)
        self.num_input_levels = n_in
        self.input_convs = nn.ModuleList([_ConvGN(self.in_channels[n_in - 1 - i], feat_channels, 1, True, groups)
                                          for i in range(self.levels)])
        self.encoder = _Encoder(self.num_layers, feat_channels, hidden, self.heads, self.levels, self.points)
        self.level_encoding = nn.Embedding(self.levels, feat_channels)
        self.lateral_convs = nn.ModuleList([_ConvGN(self.in_channels[i], feat_channels, 1, False, groups)
                                            for i in range(n_in - self.levels)])
        self.output_convs = nn.ModuleList([_ConvGN(feat_channels, feat_channels, 3, False, groups)
                                           for i in range(n_in - self.levels)])
        self.mask_feature = nn.Conv2d(feat_channels, out_channels, 1)
        self._rt = None
        self._pos = {}

    def _runtime(self, device):
        if self._rt is None or self._rt.device != device:
            self._rt = _PDRuntime(device)
            self._pos = {}
        return self._rt

    def _sine(self, k, h, w):
        key = (h, w)
        if key not in self._pos:
            out = k.new(h * w, self.feat_channels)
            k.chk(k.lib.cgg_sine_pos(k.h, _p(out), h, w, self.feat_channels, k.s()), 'cgg_sine_pos')
            self._pos[key] = out
        return self._pos[key]

    def forward(self, feats, return_tokens=False):
        """feats: backbone maps (B, C_i, H/s_i, W/s_i), highest resolution first.  Returns (mask_feature, [3 memories])."""
        f0 = feats[0]
        if not f0.is_cuda:
            raise _lib.CggError('MSDeformAttnPixelDecoderB200 runs on CUDA only (no CPU fallback)')
        k = _K(self._runtime(f0.device), tf32=(self.precision == 'tf32'))
        feats = [f.float().contiguous() for f in feats]
        B, Cc, n_in = f0.shape[0], self.feat_channels, self.num_input_levels
        # ---- input projections -> token buffer (B, S, C), level position tables (S, C)
        toks, pos, shapes = [], [], []
        for i in range(self.levels):
            f = feats[n_in - 1 - i]
            m = self.input_convs[i]
            t = _Conv1x1NCHW.apply(k, f, m.conv.weight.view(Cc, -1), m.conv.bias)
            toks.append(_GroupNorm.apply(k, t, m.gn.weight, m.gn.bias, self.groups, False))
            h, w = f.shape[-2:]
            shapes.append((h, w))
            with torch.no_grad():
                sine = self._sine(k, h, w)
            pos.append(_AddRows.apply(k, sine.view(h * w, 1, Cc), self.level_encoding.weight[i].view(1, Cc), h * w).view(h * w, Cc))
        x = torch.cat(toks, 1)
        pos = torch.cat(pos, 0)
        S = x.shape[1]
        starts = [0]
        for (h, w) in shapes[:-1]:
            starts.append(starts[-1] + h * w)
        # ---- encoder: (self_attn, norm, ffn, norm) x num_layers
        x = x.view(B * S, Cc)
        fused = not torch.is_grad_enabled()
        hs = (C.c_int * self.levels)(*[s_[0] for s_ in shapes])
        ws = (C.c_int * self.levels)(*[s_[1] for s_ in shapes])
        for layer in self.encoder.layers:
            a = layer.attentions[0]
            if fused:
                # inference: ONE projection [offsets | logits | value] of the token buffer.  (x + pos) W^T = x W^T + pos W^T: the
                # positional part is a (S, 288) table per layer, added by the epilogue as a row-periodic residual -- x + pos is
                # never materialised and x is read once instead of three times.
                n_off, n_lg = a.sampling_offsets.weight.shape[0], a.attention_weights.weight.shape[0]
                Wq = torch.cat([a.sampling_offsets.weight, a.attention_weights.weight], 0)
                Wcat = torch.cat([Wq, a.value_proj.weight], 0)
                bcat = torch.cat([a.sampling_offsets.bias, a.attention_weights.bias, a.value_proj.bias], 0)
                ptab = k.new(S, n_off + n_lg)
                k.gemm(pos, (0, Cc, 1), Wq, (0, Cc, 1), ptab, (0, n_off + n_lg, 1), S, n_off + n_lg, Cc)
                N = Wcat.shape[0]
                proj = k.new(B * S, N)
                k.gemm(x, (0, Cc, 1), Wcat, (0, Cc, 1), proj, (0, N, 1), B * S, N, Cc, bias=bcat, R=ptab,
                       sR=(0, n_off + n_lg, 1), r_mod=S, r_ncols=n_off + n_lg)
                core = k.new(B, S, Cc)
                k.chk(k.lib.cgg_ms_deform_attn(k.h, C.c_void_p(proj.data_ptr() + 4 * (n_off + n_lg)), N, _p(proj), N,
                                               C.c_void_p(proj.data_ptr() + 4 * n_off), N, _p(core), B, S, self.heads,
                                               self.levels, self.points, hs, ws, k.s()), 'cgg_ms_deform_attn')
            else:
                q = _AddRows.apply(k, x.view(B, S, Cc), pos, B).view(B * S, Cc)
                value = _Linear.apply(k, x, a.value_proj.weight, a.value_proj.bias, None, 1.0, False)
                off = _Linear.apply(k, q, a.sampling_offsets.weight, a.sampling_offsets.bias, None, 1.0, False)
                lg = _Linear.apply(k, q, a.attention_weights.weight, a.attention_weights.bias, None, 1.0, False)
                core = _MSDeformCore.apply(k, value.view(B, S, Cc), off.view(B, S, -1), lg.view(B, S, -1), shapes, self.heads,
                                           self.points)
            t = _Linear.apply(k, core.view(B * S, Cc), a.output_proj.weight, a.output_proj.bias, x, 1.0, False)
            x1 = _LayerNorm.apply(k, t, layer.norms[0].weight, layer.norms[0].bias, 1e-5)
            fc1, fc2 = layer.ffns[0].layers[0][0], layer.ffns[0].layers[1]
            f = _Linear.apply(k, x1, fc1.weight, fc1.bias, None, 1.0, True)
            t2 = _Linear.apply(k, f, fc2.weight, fc2.bias, x1, 1.0, False)
            x = _LayerNorm.apply(k, t2, layer.norms[1].weight, layer.norms[1].bias, 1e-5)
        x = x.view(B, S, Cc)
        bf16 = self.out_dtype == torch.bfloat16 and not torch.is_grad_enabled()
        outs = [_LevelToNCHW.apply(k, x, starts[i], shapes[i], bf16) for i in range(self.levels)]
        # ---- FPN top-down steps to the remaining (finer) backbone levels
        prev_tok, prev_start, prev_hw = x, starts[-1], shapes[-1]
        y = None
        for i in range(n_in - self.levels - 1, -1, -1):
            f = feats[i]
            H, W = f.shape[-2:]
            lat_m, out_m = self.lateral_convs[i], self.output_convs[i]
            lat = _Conv1x1NCHW.apply(k, f, lat_m.conv.weight.view(Cc, -1), None)
            lat = _GroupNorm.apply(k, lat, lat_m.gn.weight, lat_m.gn.bias, self.groups, False)
            y = _UpsampleAdd.apply(k, lat, prev_tok, prev_start, prev_hw, (H, W))
            W2 = out_m.conv.weight.permute(0, 2, 3, 1).reshape(Cc, 9 * Cc).contiguous()
            y = _Conv3x3.apply(k, y, W2, H, W)
            y = _GroupNorm.apply(k, y, out_m.gn.weight, out_m.gn.bias, self.groups, True)
            prev_tok, prev_start, prev_hw = y, 0, (H, W)
            if len(outs) < self.num_outs:
                outs.append(_LevelToNCHW.apply(k, y, 0, (H, W), bf16))
        if y is None:                       # no finer level: the mask features come from the last encoder level
            y, (H, W) = x[:, starts[-1]:].contiguous(), shapes[-1]
        mf = _Conv1x1ToNCHW.apply(k, y, self.mask_feature.weight.view(self.out_channels, Cc), self.mask_feature.bias, H, W)
        if bf16:
            mf = mf.to(torch.bfloat16)
        if return_tokens:
            return mf, outs[:self.num_outs], x
        return mf, outs[:self.num_outs]


def build_pixel_decoder_from_state_dict(sd, in_channels, device, precision='tf32', out_dtype=torch.float32, **kw):
    """Test / bench helper: a module carrying the tensors of `sd` (mmdet key names) on `device`."""
    feat = sd['level_encoding.weight'].shape[1]
    hidden = sd['encoder.layers.0.ffns.0.layers.0.0.weight'].shape[0]
    n_layers = 1 + max(int(k.split('.')[2]) for k in sd if k.startswith('encoder.layers.'))
    enc = dict(num_layers=n_layers, transformerlayers=dict(attn_cfgs=dict(num_heads=8, num_levels=3, num_points=4),
                                                           ffn_cfgs=dict(feedforward_channels=hidden)))
    m = MSDeformAttnPixelDecoderB200(in_channels=in_channels, feat_channels=feat, out_channels=sd['mask_feature.weight'].shape[0],
                                     encoder=enc, norm_cfg=dict(type='GN', num_groups=32), precision=precision,
                                     out_dtype=out_dtype, **kw)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.to(device)
