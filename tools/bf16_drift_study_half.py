"""CPU study (build container), companion of bf16_drift_study.py: the layer chain (weights, activations, FFN hidden) rounded to IEEE half instead of bf16"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bf16_drift_study as S
from bf16_drift_study import *
torch.set_num_threads(8)
# monkeypatch r(): flags in HALF set round to fp16 instead of bf16
HALF = {'w', 'act', 'ffn'}
orig_forward = S.forward
def r2(x, on, half=False):
    if not on: return x
    return x.half().float() if half else x.bfloat16().float()
import types, inspect
src = inspect.getsource(S.mha) + inspect.getsource(S.forward)
src = src.replace("r(q_in, 'act' in f)", "r2(q_in, 'act' in f, True)").replace("r(in_w[:C], 'w' in f)", "r2(in_w[:C], 'w' in f, True)")
src = src.replace("r(k_in, 'act' in f)", "r2(k_in, 'act' in f, True)").replace("r(in_w[C:2 * C], 'w' in f)", "r2(in_w[C:2 * C], 'w' in f, True)")
src = src.replace("r(v_in, 'act' in f)", "r2(v_in, 'act' in f, True)").replace("r(in_w[2 * C:], 'w' in f)", "r2(in_w[2 * C:], 'w' in f, True)")
src = src.replace("r(o, 'act' in f)", "r2(o, 'act' in f, True)").replace("r(out_w, 'w' in f)", "r2(out_w, 'w' in f, True)")
src = src.replace("r(x, 'act' in f)", "r2(x, 'act' in f, True)").replace("r(sd[p + 'ffns.0.layers.0.0.weight'], 'w' in f)", "r2(sd[p + 'ffns.0.layers.0.0.weight'], 'w' in f, True)")
src = src.replace("r(hdn, 'ffn' in f)", "r2(hdn, 'ffn' in f, True)").replace("r(sd[p + 'ffns.0.layers.1.weight'], 'w' in f)", "r2(sd[p + 'ffns.0.layers.1.weight'], 'w' in f, True)")
ns = dict(S.__dict__); ns['r2'] = r2
exec(src, ns)
fwd16 = ns['forward']
for (Q, B, H, W, ps, iseed) in [(100, 1, 1024, 1024, 0, 0), (100, 1, 512, 512, 33, 9), (100, 2, 512, 512, 5, 3), (100, 2, 768, 768, 5, 3)]:
    sd = synth.make_params(seed=ps, num_queries=Q, perturb=(ps != 0))
    mf, mems = synth.make_inputs(iseed, B, H, W)
    mf, mems = mf.bfloat16().float(), [m.bfloat16().float() for m in mems]
    with torch.no_grad():
        ref = S.forward(sd, mf, mems, set())
        base = {'kvw', 'kv', 'p', 'me', 'q'}
        report('base (chain exact)', ref, S.forward(sd, mf, mems, base))
        report('base + chain fp16', ref, fwd16(sd, mf, mems, base | {'w','act','ffn'}))
        report('base + chain bf16', ref, S.forward(sd, mf, mems, base | {'w','act','ffn'}))
    print()
