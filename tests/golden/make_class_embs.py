"""Writes tests/golden/class_embs.npz: the class-embedding buffers the UNMODIFIED reference head builds from its own
json files (open_set/models/mask2former_head.py:202-217), for the two configs the tests use.  Run in the build
container only (needs /root/reference):

    python tests/golden/make_class_embs.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
R = ref_shim.REF_ROOT

if __name__ == '__main__':
    out = {}
    # configs/instance/coco_b48n17.py: 48 known classes (+1 background row of zeros)
    h = ref_shim.build_reference_head(num_queries=8, known_file=R + '/datasets/unknown/known_65.txt',
                                      unknown_file=R + '/datasets/unknown/unknown_17.txt')
    out['coco_instance_48'] = h.class_embs.numpy()
    # configs/openset_panoptic/coco_panoptic_p20.py: 64 known things + 53 stuff (+1)
    h = ref_shim.build_reference_head(num_queries=8, num_known=64, num_stuff=53,
                                      unknown_file=R + '/datasets/unknown/unknown_p20.txt',
                                      class_to_emb_file=R + '/datasets/embeddings/coco_panoptic_class_with_bert_emb.json')
    out['coco_panoptic_p20'] = h.class_embs.numpy()
    for k, v in out.items():
        print(k, v.shape, 'zero rows:', int((np.abs(v).sum(1) == 0).sum()), 'row norm %.2f' % float(np.linalg.norm(v[0])))
    np.savez_compressed(os.path.join(HERE, 'class_embs.npz'), **out)
