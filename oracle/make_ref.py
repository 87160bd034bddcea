"""Recipe for `oracle/_ref/`: a travelling copy of the reference source files the hot path and its losses consist of plus the
data files its constructor reads, taken unmodified from /root/reference at build time (this container only).

`oracle/_ref/` is git-ignored -- the reference's sources never enter this repository's history -- but it is not
gpurun-ignored, so the copy rides to the GPU box, where `bench.py --impl reference` and the `cpu_baseline` leg time the
reference's OWN implementation (kind "reference") through `oracle/ref_shim.py` instead of the oracle port.  Run by
`__graft_entry__.build()` whenever /root/reference is present.

    python oracle/make_ref.py
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('CGG_REFERENCE_SRC', '/root/reference')
DST = os.path.join(HERE, '_ref')
FILES = [
    'open_set/models/mask2former_head.py',            # the path: forward / forward_head (:711-849)
    'open_set/models/losses/grounding_loss.py',       # grounding_loss (:9-77)
    'open_set/models/utils/bert_embeddings.py',       # BertEmbeddings (:4-13)
    'open_set/utils/eval/inference.py',               # imported by the head file (:27)
    'open_set/models/losses/cross_entropy_loss.py',   # loss_cls / loss_cls_emb / loss_mask of loss_single (:62-199)
    'open_set/assigners/mask_hungarian_assigner.py',  # the Hungarian assignment of _get_target_single (:47-146)
    'datasets/embeddings/coco_class_with_bert_emb.json',      # class_embs of the instance config
    'datasets/unknown/known_65.txt',
    'datasets/unknown/unknown_17.txt',
]


def make():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return None
    manifest = []
    for f in FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), d)
        manifest.append('%s  %s' % (hashlib.sha256(open(d, 'rb').read()).hexdigest()[:16], f))
    with open(os.path.join(DST, 'MANIFEST.txt'), 'w') as fh:
        fh.write('verbatim copies from %s (sha256[:16], path)\n' % SRC + '\n'.join(manifest) + '\n')
    return DST


if __name__ == '__main__':
    print(make())
