import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, bench
from yad_b200 import _lib
torch.set_grad_enabled(False)
dev = torch.device('cuda', 0)
model, _ = bench.build_model(dev, 'bf16', deploy=True)
x = bench.synth_clips_device(512, dev, 1000)
model(x, combine_scales=True); torch.cuda.synchronize()
eng = model._engine(); B, _, L = x.shape; plan = eng._plan((B, L)); T = eng.frames(L)
xs = plan['xs']; cur = plan['c2']; fx = plan['fstem_cols']; xb = plan['xs_bf16']
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
s = eng._stream
for nint in [int(v) for v in os.environ.get("NINTS", "84,92,96,100,104,108,112").split(",")]:
    ms = t(lambda: _lib.check(eng.lib.yad_conv_stem_fused(xb.data_ptr(), xb.shape[2], B, 32, T, eng.fstem_w.data_ptr(), eng.fstem_bias.data_ptr(), cur.data_ptr(), cur.shape[2], cur.shape[1], nint, s()), 'f'))
    print('n_int', nint, 'fused main %.3f ms' % ms)
ms = t(lambda: _lib.check(eng.lib.yad_conv_stem_fused_fixup(xs.data_ptr(), B, 32, T, eng.fstem_wvar.data_ptr(), eng.fstem_bias.data_ptr(), fx[0], fx[1], fx[2], cur.data_ptr(), cur.shape[2], cur.shape[1], s()), 'x'))
print('fixup %.3f ms' % ms)
mel = plan['mel']
ms = t(lambda: _lib.check(eng.lib.yad_frontend_finish(mel.data_ptr(), B, T, eng.dct.data_ptr(), 80.0, 1, xs.data_ptr(), 0, 0, 0, s()), 'fin'))
print('finish %.3f ms' % ms)
