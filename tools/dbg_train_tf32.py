"""Debug: per-output and per-parameter errors of the tf32 training step against the fp32 oracle (and torch TF32)."""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import cgg_oracle as O
from cgg_b200.head import build_head_from_state_dict
from cgg_b200.grounding import grounding_loss
from cgg_b200.train import decoder_forward_train
import test_gpu_train as T

DEV = 'cuda'
Q, B, H, W, ncls1 = [(24, 2, 128, 160, 49), (40, 1, 160, 128, 118)][int(sys.argv[1]) if len(sys.argv) > 1 else 0]
tp = sys.argv[2] if len(sys.argv) > 2 else 'tf32'
sd, mf, mems, probes, cap, cap_mask = T._setup(Q, B, H, W, ncls1)
sd_o = {k: v.clone().requires_grad_(k != 'class_embs') for k, v in sd.items()}
ref = O.decoder_forward(sd_o, mf.clone(), [m.clone() for m in mems])
loss_o = T._loss_from_outputs(ref['cls'], ref['emb'], ref['mask'], probes, cap, cap_mask, lambda e, c, m: O.grounding_loss(e, c, m, 10.0, 2.0))
loss_o.backward()
head = build_head_from_state_dict(sd, Q, ncls1, 'fp32', DEV, train_precision=tp).train()
want_bits = [O.pack_mask_bits(ref['masked'][j].detach()) for j in range(9)]
forced = [(want_bits[j].to(DEV), ref['masked'][j].detach().all(-1).to(torch.uint8).to(DEV)) for j in range(9)]
cls, emb, mask = decoder_forward_train(head, mf.to(DEV), [m.to(DEV) for m in mems], forced_attn_masks=forced)
rel = lambda a, b: float((a.detach().cpu() - b.detach()).abs().max()) / float(b.detach().abs().max())
for j in range(10):
    print('call %d: cls %.2e emb %.2e mask %.2e' % (j, rel(cls[j], ref['cls'][j]), rel(emb[j], ref['emb'][j]), rel(mask[j], ref['mask'][j])))
probes_d = {k: [t.to(DEV) for t in v] for k, v in probes.items()}
loss = T._loss_from_outputs(cls, emb, mask, probes_d, cap.to(DEV), cap_mask.to(DEV), lambda e, c, m: grounding_loss(e, c, m, 10.0, 2.0))
loss.backward()
print('loss', float(loss), float(loss_o))
l2 = lambda a, b: float((a.detach().cpu().double() - b.detach().double()).norm() / b.detach().double().norm())
errs = sorted(((rel(p.grad, sd_o[k].grad), k) for k, p in head.named_parameters()), reverse=True)
for e, k in errs[:4]:
    print('%.2e %s' % (e, k))
errs = sorted(((l2(p.grad, sd_o[k].grad), k) for k, p in head.named_parameters()), reverse=True)
for e, k in errs[:6]:
    print('L2 %.2e %s' % (e, k))
# torch TF32
torch.backends.cuda.matmul.allow_tf32 = True
sd_t = {k: v.clone().to(DEV).requires_grad_(k != 'class_embs') for k, v in sd.items()}
ref_t = O.decoder_forward(sd_t, mf.to(DEV), [m.to(DEV) for m in mems], forced_masked=[r.detach().to(DEV) for r in ref['masked']])
loss_t = T._loss_from_outputs(ref_t['cls'], ref_t['emb'], ref_t['mask'], probes_d, cap.to(DEV), cap_mask.to(DEV), lambda e, c, m: O.grounding_loss(e, c, m, 10.0, 2.0))
loss_t.backward()
torch.backends.cuda.matmul.allow_tf32 = False
print('torch TF32: loss', float(loss_t))
for j in (0, 5, 9):
    print('torch TF32 call %d: cls %.2e emb %.2e mask %.2e' % (j, rel(ref_t['cls'][j], ref['cls'][j]), rel(ref_t['emb'][j], ref['emb'][j]), rel(ref_t['mask'][j], ref['mask'][j])))
errs = sorted(((rel(sd_t[k].grad, sd_o[k].grad), k) for k in dict(head.named_parameters())), reverse=True)
for e, k in errs[:3]:
    print('torch TF32 %.2e %s' % (e, k))
errs = sorted(((l2(sd_t[k].grad, sd_o[k].grad), k) for k in dict(head.named_parameters())), reverse=True)
for e, k in errs[:4]:
    print('torch TF32 L2 %.2e %s' % (e, k))
