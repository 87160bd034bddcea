"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the reference's own ``open_set/models/mask2former_head.py`` and
``open_set/models/losses/grounding_loss.py`` *unmodified* from /root/reference by
installing tiny stand-ins for the third-party packages that are absent in this
image (mmcv-full 1.7.1, mmdet 2.28.2, clip).  /root/reference exists only in the
build container, never on the GPU box, so this module is used solely by

  * ``tests/golden/make_golden.py``  (writes the committed golden vectors), and
  * ``tests/test_oracle.py``         (validates ``oracle/cgg_oracle.py`` against the
                                      verbatim reference; skipped when the reference
                                      tree is absent).

The stand-ins restate the mmcv/mmdet semantics that the head relies on
(SURVEY.md Appendix A/B):
  mmcv.cnn.bricks.transformer.MultiheadAttention  -> wrapper around the real
      torch.nn.MultiheadAttention: q += query_pos, k += key_pos, value gets no pos,
      identity + attn(...)[0]
  BaseTransformerLayer (operation_order cross_attn,norm,self_attn,norm,ffn,norm)
  FFN  Sequential(Sequential(Linear,ReLU,Dropout), Linear, Dropout) + identity
  mmdet SinePositionalEncoding(num_feats=128, normalize=True)
"""
import importlib
import json
import math
import os
import sys
import types

import torch
import torch.nn as nn

def _find_ref_root():
    """/root/reference in the build container; on the GPU box the verbatim copy that oracle/make_ref.py left in the
    git-ignored oracle/_ref/ (it travels with the snapshot)."""
    env = os.environ.get('CGG_REFERENCE_ROOT')
    if env:
        return env
    for cand in ('/root/reference', os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')):
        if os.path.isfile(os.path.join(cand, 'open_set/models/mask2former_head.py')):
            return cand
    return '/root/reference'


REF_ROOT = _find_ref_root()


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, 'open_set/models/mask2former_head.py'))


class AttrDict(dict):
    """dict with attribute access, standing in for mmcv.ConfigDict."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(obj):
        if isinstance(obj, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in obj.items()})
        if isinstance(obj, (list, tuple)):
            return type(obj)(AttrDict.wrap(v) for v in obj)
        return obj

    def __deepcopy__(self, memo):
        import copy
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


# --------------------------------------------------------------------------- mmcv bricks
class _MHAWrapper(nn.Module):
    """mmcv 1.7.1 cnn/bricks/transformer.py MultiheadAttention (batch_first=False,
    all dropouts 0)."""

    def __init__(self, embed_dims, num_heads, **kw):
        super().__init__()
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, 0.0)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_pos=None, attn_mask=None, key_padding_mask=None, **kw):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        key_padding_mask=key_padding_mask)[0]
        return identity + out


class _FFN(nn.Module):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, **kw):
        super().__init__()
        assert num_fcs == 2
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True),
                          nn.Dropout(0.0)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.0))

    def forward(self, x, identity=None):
        out = self.layers(x)
        if identity is None:
            identity = x
        return identity + out


class _DecoderLayer(nn.Module):
    def __init__(self, attn_cfgs, ffn_cfgs, operation_order, feedforward_channels=None, **kw):
        super().__init__()
        self.operation_order = tuple(operation_order)
        n_attn = sum(1 for o in operation_order if o.endswith('attn'))
        self.attentions = nn.ModuleList(
            [_MHAWrapper(attn_cfgs['embed_dims'], attn_cfgs['num_heads']) for _ in range(n_attn)])
        self.embed_dims = attn_cfgs['embed_dims']
        ffn = dict(ffn_cfgs)
        self.ffns = nn.ModuleList([_FFN(ffn.get('embed_dims', self.embed_dims),
                                        ffn.get('feedforward_channels', feedforward_channels))])
        self.norms = nn.ModuleList(
            [nn.LayerNorm(self.embed_dims) for o in operation_order if o == 'norm'])

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None,
                attn_masks=None, query_key_padding_mask=None, key_padding_mask=None, **kw):
        ai = ni = fi = 0
        for op in self.operation_order:
            if op == 'self_attn':
                query = self.attentions[ai](query, query, query, None, query_pos=query_pos,
                                            key_pos=query_pos, attn_mask=attn_masks[ai],
                                            key_padding_mask=query_key_padding_mask)
                ai += 1
            elif op == 'cross_attn':
                query = self.attentions[ai](query, key, value, None, query_pos=query_pos,
                                            key_pos=key_pos, attn_mask=attn_masks[ai],
                                            key_padding_mask=key_padding_mask)
                ai += 1
            elif op == 'norm':
                query = self.norms[ni](query)
                ni += 1
            elif op == 'ffn':
                query = self.ffns[fi](query, None)
                fi += 1
        return query


class _Decoder(nn.Module):
    def __init__(self, transformerlayers, num_layers, **kw):
        super().__init__()
        tl = dict(transformerlayers)
        tl.pop('type', None)
        self.layers = nn.ModuleList([_DecoderLayer(**tl) for _ in range(num_layers)])
        self.embed_dims = self.layers[0].embed_dims
        self.post_norm = nn.LayerNorm(self.embed_dims)


class _SinePE(nn.Module):
    """mmdet 2.28.2 SinePositionalEncoding."""

    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * math.pi,
                 eps=1e-6, offset=0., **kw):
        super().__init__()
        self.num_feats, self.temperature, self.normalize = num_feats, temperature, normalize
        self.scale, self.eps, self.offset = scale, eps, offset

    def forward(self, mask):
        mask = mask.to(torch.int)
        not_mask = 1 - mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            y_embed = (y_embed + self.offset) / (y_embed[:, -1:, :] + self.eps) * self.scale
            x_embed = (x_embed + self.offset) / (x_embed[:, :, -1:] + self.eps) * self.scale
        dim_t = torch.arange(self.num_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.num_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        B, H, W = mask.size()
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


class _PassThroughPixelDecoder(nn.Module):
    """The pixel decoder is the step BEFORE the path: the stub
    hands (mask_features, [mem32, mem16, mem8]) through unchanged."""

    def forward(self, feats):
        return feats[0], feats[1]

    def init_weights(self):
        pass


class _Registry:
    def __init__(self):
        self.d = {}

    def register_module(self, name=None, **kw):
        def deco(cls):
            self.d[name or cls.__name__] = cls
            return cls
        return deco


class _StubLoss(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.cfg = kw


def _install():
    if 'mmcv' in sys.modules and getattr(sys.modules['mmcv'], '_cgg_shim', False):
        return
    R = REF_ROOT

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    HEADS, LOSSES = _Registry(), _Registry()

    class FileClient:
        def get_text(self, path):
            with open(path) as f:
                return f.read()

    def load(path):
        with open(path) as f:
            return json.load(f)

    def build_loss(cfg):
        cfg = dict(cfg)
        t = cfg.pop('type')
        if t in LOSSES.d:
            return LOSSES.d[t](**cfg)
        return _StubLoss(**cfg)

    def build_head(cfg):
        return nn.Identity()

    def multi_apply(func, *args, **kwargs):
        from functools import partial
        pfunc = partial(func, **kwargs) if kwargs else func
        return tuple(map(list, zip(*map(pfunc, *args))))

    def reduce_mean(t):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return t
        t = t.clone()
        dist.all_reduce(t.div_(dist.get_world_size()))
        return t

    def get_dist_info():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def force_fp32(*a, **k):
        def deco(f):
            return f
        return deco

    class _Base(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()

    class AnchorFreeHead(_Base):
        pass

    class MaskFormerHead(AnchorFreeHead):
        pass

    # ---- the training-time bricks of loss_single (row f2): mmcv / mmdet entry points restated in matching_oracle.py
    from . import matching_oracle as MO
    BBOX_ASSIGNERS = _Registry()

    class _Cost:
        def __init__(self, weight=1.0, **kw):
            self.weight, self.kw = weight, kw

    class ClassificationCost(_Cost):
        def __call__(self, cls_pred, gt_labels):
            return MO.classification_cost(cls_pred, gt_labels, self.weight)

    class CrossEntropyLossCost(_Cost):
        def __call__(self, pred, gt):
            assert self.kw.get('use_sigmoid', True)
            return MO.cross_entropy_loss_cost(pred, gt, self.weight)

    class DiceCost(_Cost):
        def __call__(self, pred, gt):
            return MO.dice_cost(pred, gt, self.weight, pred_act=self.kw.get('pred_act', False), eps=self.kw.get('eps', 1e-3),
                                naive_dice=self.kw.get('naive_dice', True))

    COSTS = dict(ClassificationCost=ClassificationCost, CrossEntropyLossCost=CrossEntropyLossCost, DiceCost=DiceCost)

    def build_match_cost(cfg):
        cfg = dict(cfg)
        return COSTS[cfg.pop('type').replace('ClassficationCost', 'ClassificationCost')](**cfg)

    class AssignResult:
        def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
            self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels

    class _Sampled:
        pass

    class MaskPseudoSampler:
        """mmdet/core/bbox/samplers/mask_pseudo_sampler.py"""
        def sample(self, assign_result, masks, gt_masks, **kw):
            r = _Sampled()
            r.pos_inds = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).squeeze(-1).unique()
            r.neg_inds = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).squeeze(-1).unique()
            r.pos_assigned_gt_inds = assign_result.gt_inds[r.pos_inds] - 1
            return r

    def build_assigner(cfg):
        importlib.import_module('open_set.assigners.mask_hungarian_assigner')     # registers the reference's own class
        cfg = dict(cfg)
        return BBOX_ASSIGNERS.d[cfg.pop('type')](**cfg)

    def build_sampler(cfg, **kw):
        assert cfg['type'] == 'MaskPseudoSampler'
        return MaskPseudoSampler()

    class DiceLoss(nn.Module):
        """mmdet/models/losses/dice_loss.py (use_sigmoid + activate, naive_dice)"""
        def __init__(self, use_sigmoid=True, activate=True, reduction='mean', naive_dice=False, loss_weight=1.0, eps=1e-3):
            super().__init__()
            self.ok = use_sigmoid and activate and naive_dice and reduction == 'mean'      # the form the configs use
            self.loss_weight, self.eps = loss_weight, eps

        def forward(self, pred, target, weight=None, reduction_override=None, avg_factor=None):
            assert self.ok, 'only DiceLoss(use_sigmoid, activate, naive_dice, reduction=mean) is restated'
            return MO.dice_loss(pred, target, avg_factor, self.loss_weight, self.eps)

    LOSSES.d['DiceLoss'] = DiceLoss

    m = mod('mmcv', FileClient=FileClient, load=load, _cgg_shim=True)
    m.__path__ = []
    mod('mmcv.cnn', Conv2d=nn.Conv2d,
        build_plugin_layer=lambda cfg, *a, **k: ('pixel_decoder', _PassThroughPixelDecoder()),
        caffe2_xavier_init=lambda *a, **k: None).__path__ = []
    mod('mmcv.cnn.bricks').__path__ = []
    mod('mmcv.cnn.bricks.transformer',
        build_positional_encoding=lambda cfg: _SinePE(**{k: v for k, v in cfg.items() if k != 'type'}),
        build_transformer_layer_sequence=lambda cfg: _Decoder(**{k: v for k, v in cfg.items() if k != 'type'}))
    mod('mmcv.ops', point_sample=MO.point_sample, RoIPool=None)
    mod('mmcv.runner', ModuleList=nn.ModuleList, force_fp32=force_fp32, get_dist_info=get_dist_info)
    mod('mmcv.parallel', collate=None, scatter=None)
    mod('mmdet').__path__ = []
    mod('mmdet.core', build_assigner=build_assigner, build_sampler=build_sampler, multi_apply=multi_apply,
        reduce_mean=reduce_mean).__path__ = []
    mod('mmdet.core.bbox').__path__ = []
    mod('mmdet.core.bbox.builder', BBOX_ASSIGNERS=BBOX_ASSIGNERS)
    mod('mmdet.core.bbox.match_costs').__path__ = []
    mod('mmdet.core.bbox.match_costs.builder', build_match_cost=build_match_cost)
    mod('mmdet.core.bbox.assigners').__path__ = []
    mod('mmdet.core.bbox.assigners.assign_result', AssignResult=AssignResult)
    mod('mmdet.core.bbox.assigners.base_assigner', BaseAssigner=object)
    mod('mmdet.models.losses').__path__ = []
    mod('mmdet.models.losses.utils', weight_reduce_loss=MO.weight_reduce_loss)
    mod('mmdet.datasets', replace_ImageToTensor=None).__path__ = []
    mod('mmdet.datasets.pipelines', Compose=None)
    mod('mmdet.models').__path__ = []
    mod('mmdet.models.utils', preprocess_panoptic_gt=None,
        get_uncertain_point_coords_with_randomness=lambda mask_pred, labels, n, o, i:
        MO.get_uncertain_point_coords_with_randomness(mask_pred, n, o, i))
    mod('mmdet.models.builder', HEADS=HEADS, LOSSES=LOSSES, build_loss=build_loss, build_head=build_head)
    mod('mmdet.models.dense_heads').__path__ = []
    mod('mmdet.models.dense_heads.anchor_free_head', AnchorFreeHead=AnchorFreeHead)
    mod('mmdet.models.dense_heads.maskformer_head', MaskFormerHead=MaskFormerHead)
    mod('clip')
    # namespace stubs so the reference's own __init__.py files (pycocotools, spacy, ...) never run
    for name, path in [('open_set', R + '/open_set'), ('open_set.models', R + '/open_set/models'),
                       ('open_set.models.utils', R + '/open_set/models/utils'),
                       ('open_set.models.losses', R + '/open_set/models/losses'),
                       ('open_set.assigners', R + '/open_set/assigners'),
                       ('open_set.models.transformers', R + '/open_set/models/transformers'),
                       ('open_set.utils', R + '/open_set/utils'),
                       ('open_set.utils.eval', R + '/open_set/utils/eval')]:
        pm = types.ModuleType(name)
        pm.__path__ = [path]
        sys.modules[name] = pm


def load_reference_modules():
    """Returns (head_module, grounding_loss_module): the unmodified reference files."""
    if not reference_available():
        raise RuntimeError('reference tree not present at ' + REF_ROOT)
    _install()
    gl = importlib.import_module('open_set.models.losses.grounding_loss')
    # the config's 'CrossEntropyLoss' is mmdet's; the reference carries its own copy of that code (CrossEntropyLossOpen,
    # open_set/models/losses/cross_entropy_loss.py:257-): that copy computes loss_cls / loss_cls_emb / loss_mask here
    if os.path.isfile(os.path.join(REF_ROOT, 'open_set/models/losses/cross_entropy_loss.py')):
        ce = importlib.import_module('open_set.models.losses.cross_entropy_loss')
        sys.modules['mmdet.models.builder'].LOSSES.d['CrossEntropyLoss'] = ce.CrossEntropyLossOpen
    head = importlib.import_module('open_set.models.mask2former_head')
    return head, gl


def head_cfg(num_queries=100, num_layers=9, num_known=48, num_stuff=0, embed=256, heads=8,
             ffn=2048, known_file=None, unknown_file=None, class_to_emb_file=None):
    """Head kwargs following configs/instance/coco_b48n17.py:28-154 (decoder part)."""
    cfg = dict(
        in_channels=[256, 512, 1024, 2048], feat_channels=embed, out_channels=embed,
        num_things_classes=num_known, num_stuff_classes=num_stuff, num_queries=num_queries,
        num_transformer_feat_level=3,
        pixel_decoder=dict(type='Stub', encoder=dict(transformerlayers=dict(attn_cfgs=dict(num_levels=3)))),
        enforce_decoder_input_project=False,
        positional_encoding=dict(type='SinePositionalEncoding', num_feats=embed // 2, normalize=True),
        transformer_decoder=dict(
            type='DetrTransformerDecoder', return_intermediate=True, num_layers=num_layers,
            transformerlayers=dict(
                type='DetrTransformerDecoderLayer',
                attn_cfgs=dict(type='MultiheadAttention', embed_dims=embed, num_heads=heads),
                ffn_cfgs=dict(embed_dims=embed, feedforward_channels=ffn, num_fcs=2),
                feedforward_channels=ffn,
                operation_order=('cross_attn', 'norm', 'self_attn', 'norm', 'ffn', 'norm'))),
        loss_cls=dict(type='CrossEntropyLoss', class_weight=[1.0] * (num_known + num_stuff) + [0.1]),
        loss_mask=dict(type='CrossEntropyLoss'), loss_dice=dict(type='DiceLoss'),
        loss_grounding=dict(type='GroundingLoss', loss_weight=2.0),
        use_class_emb=True,
        class_to_emb_file=class_to_emb_file or REF_ROOT + '/datasets/embeddings/coco_class_with_bert_emb.json',
        known_file=known_file, unknown_file=unknown_file,
        softmax_temperature=10, pred_emb_norm=False, text_emb_norm=True)
    return AttrDict.wrap(cfg)


def load_caption_modules():
    """(caption_tranformer module, inference module) of the unmodified reference: CaptionTransformer and beam_search."""
    load_reference_modules()
    ct = importlib.import_module('open_set.models.transformers.caption_tranformer')
    inf = importlib.import_module('open_set.utils.eval.inference')
    return ct, inf


def train_cfg(num_points=12544):
    """configs/openset_panoptic/coco_panoptic_p20.py:163-175"""
    return AttrDict.wrap(dict(
        num_points=num_points, oversample_ratio=3.0, importance_sample_ratio=0.75,
        assigner=dict(type='MaskHungarianAssignerOpen', cls_cost=dict(type='ClassificationCost', weight=0.0),
                      cls_emb_cost=dict(type='ClassificationCost', weight=2.0),
                      mask_cost=dict(type='CrossEntropyLossCost', weight=5.0, use_sigmoid=True),
                      dice_cost=dict(type='DiceCost', weight=5.0, pred_act=True, eps=1.0)),
        sampler=dict(type='MaskPseudoSampler')))


def build_reference_head(with_losses=False, num_points=12544, **kw):
    head_mod, _ = load_reference_modules()
    cfg = head_cfg(**kw)
    if with_losses:         # the loss / assignment configuration of coco_panoptic_p20.py:111-139, :163-175
        ncls = kw.get('num_known', 48) + kw.get('num_stuff', 0)
        cw = [1.0] * ncls + [0.1]
        cfg.update(AttrDict.wrap(dict(
            loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=0.0, reduction='mean', class_weight=cw),
            loss_cls_emb=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=2.0, reduction='mean', class_weight=cw),
            loss_mask=dict(type='CrossEntropyLoss', use_sigmoid=True, reduction='mean', loss_weight=5.0),
            loss_dice=dict(type='DiceLoss', use_sigmoid=True, activate=True, reduction='mean', naive_dice=True, eps=1.0,
                           loss_weight=5.0),
            train_cfg=train_cfg(num_points))))
    head = head_mod.Mask2FormerHeadOpen(**cfg)
    head.init_weights()
    return head.eval()


def run_reference_head(head, mask_features, memories):
    """memories: [mem32, mem16, mem8] (low -> high resolution), as the pixel decoder returns."""
    B = mask_features.shape[0]
    with torch.no_grad():
        return head((mask_features, list(memories)), [dict() for _ in range(B)])
