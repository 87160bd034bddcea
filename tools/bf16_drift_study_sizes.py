"""CPU study (build container), companion of bf16_drift_study.py: the same flags at 1024^2 / 512^2 (the error of every flag shrinks with the key count)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bf16_drift_study import *
torch.set_num_threads(8)
for (Q, B, H, W, ps, iseed) in [(100, 1, 1024, 1024, 0, 0), (100, 1, 512, 512, 33, 9)]:
    sd = synth.make_params(seed=ps, num_queries=Q, perturb=(ps != 0))
    mf, mems = synth.make_inputs(iseed, B, H, W)
    mf, mems = mf.bfloat16().float(), [m.bfloat16().float() for m in mems]
    with torch.no_grad():
        ref = forward(sd, mf, mems, set())
        allf = ['w', 'act', 'ffn', 'kvw', 'kv', 'q', 'p', 'me']
        for fl in allf:
            report(fl, ref, forward(sd, mf, mems, {fl}))
        report('all', ref, forward(sd, mf, mems, set(allf)))
        report('kvw,kv,p,me', ref, forward(sd, mf, mems, {'kvw', 'kv', 'p', 'me'}))
        report('kvw,kv,p,me,q', ref, forward(sd, mf, mems, {'kvw', 'kv', 'p', 'me','q'}))
        report('kv,me', ref, forward(sd, mf, mems, {'kv', 'me'}))
    print()
