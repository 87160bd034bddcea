"""Two eager training steps of bench.py's train leg (BASELINE configs[3]: Q=200, 118 class rows, B=2, 1024x1024) --
the `ncu` target for the training step's launch list and for the tf32 GEMM captures.
  python tools/one_train_step.py [fp32|tf32] [steps] [reducer]     (reducer: gradients go to a GradReducer's buckets, which
  turns on the fused weight-gradient path -- the form the bench's train leg runs)"""
import sys
import torch
sys.path.insert(0, '.')
from cgg_b200 import synth
from cgg_b200.head import build_head_from_state_dict
from cgg_b200.grounding import grounding_loss, similarity

tp = sys.argv[1] if len(sys.argv) > 1 else 'tf32'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
Q, B, ncls1, H, W = 200, 2, 118, 1024, 1024
dev = torch.device('cuda', 0)
sd = synth.make_params(seed=0, num_queries=Q, num_classes_p1=ncls1)
head = build_head_from_state_dict(sd, Q, ncls1, 'fp32', dev, train_precision=tp).train()
mf, mems = synth.make_inputs(100, B, H, W)
mf, mems = mf.to(dev), [m.to(dev) for m in mems]
ids, cap_mask, table, lw, lb = synth.make_captions(0, B)
cap = head.extract_word_embeddings(table.to(dev), lw.to(dev), lb.to(dev), ids.to(dev))
cap_mask = cap_mask.to(dev)
g = torch.Generator().manual_seed(0)
labels = torch.randint(0, ncls1, (B, Q), generator=g).to(dev)
targets = (torch.rand((B, Q, H // 4, W // 4), generator=g) > 0.5).to(dev).float()
red = None
if len(sys.argv) > 3 and sys.argv[3] == 'reducer':
    from cgg_b200.train import GradReducer
    red = GradReducer(head.parameters())
for _ in range(steps):
    if red is not None:
        red.zero()
    else:
        for p in head.parameters():
            p.grad = None
    cls, emb, mask = head.decoder_forward_auto(mf, mems)
    loss = 0.0
    for j in range(len(cls)):
        loss = loss + grounding_loss(emb[j], cap, cap_mask, 10.0, 2.0)
        loss = loss + torch.nn.functional.cross_entropy(similarity(emb[j].reshape(B * Q, -1), head.class_embs, 0.1), labels.reshape(-1))
        loss = loss + torch.nn.functional.binary_cross_entropy_with_logits(mask[j], targets)
    loss.backward()
    if red is not None:
        red.finish()
torch.cuda.synchronize()
print('loss', float(loss))
