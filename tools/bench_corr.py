"""Micro-benchmark of yad_corr_tf32 on the layer1 shape: forward vs data gradient, accumulate on/off."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import yad_b200
from yad_b200 import _lib
from yad_b200._lib import CorrDesc
lib = _lib.init(0); dev = torch.device("cuda", 0)
B, H, W, Cc = 32, 8, 240, 64
x = torch.randn(B, H, W, Cc, device=dev); w = torch.randn(64, 9 * 64, device=dev) * 0.05
out = torch.zeros(B, H, W, Cc, device=dev)
arr = lambda v: (C.c_int32 * len(v))(*v)
taps = [(a, c) for a in range(3) for c in range(3)]
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(name, dh, dw, acc, scale=1.0, dense=True):
    xs = x * scale
    d = CorrDesc(B=B, H=H, W=W, Cin=64, ld_in=64, Ho=H, Wo=W, Cout=64, ld_out=64, sh=1, sw=1, out_sw=0 if dense else 1, out_sh=0 if dense else W,
                 out_sb=0 if dense else H * W, n_taps=9, act=0, accumulate=acc, whole_rows=0)
    k = arr([(a * 3 + c) * 64 for a, c in taps])
    f = lambda: _lib.check(lib.yad_corr_tf32(C.byref(d), arr(dh), arr(dw), k, xs.data_ptr(), w.data_ptr(), 64, 576, 0, out.data_ptr(), st()), name)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / 20 * 1000:8.1f} us")
run("fwd taps, overwrite", [a - 1 for a, _ in taps], [c - 1 for _, c in taps], 0)
run("fwd taps, accumulate", [a - 1 for a, _ in taps], [c - 1 for _, c in taps], 1)
run("dgrad taps, accumulate", [1 - a for a, _ in taps], [1 - c for _, c in taps], 1)
run("dgrad taps, accumulate, strided desc", [1 - a for a, _ in taps], [1 - c for _, c in taps], 1, dense=False)
run("fwd, tiny values 1e-20", [a - 1 for a, _ in taps], [c - 1 for _, c in taps], 0, scale=1e-20)
run("fwd, denormal values 1e-40", [a - 1 for a, _ in taps], [c - 1 for _, c in taps], 0, scale=1e-40)
run("fwd, zeros", [a - 1 for a, _ in taps], [c - 1 for _, c in taps], 0, scale=0.0)
