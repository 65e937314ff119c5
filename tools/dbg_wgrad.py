import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import yad_b200
from yad_b200 import _lib
from yad_b200._lib import CorrDesc
lib = _lib.init(0)
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
def run(B, H, W, Cin, Cout, k, s, p):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev)
    w = torch.randn(Cout, Cin, *k, generator=g).to(dev).requires_grad_(True)
    y = F.conv2d(x, w, None, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g).to(dev)
    gw, = torch.autograd.grad(y, (w,), dy)
    Ho, Wo = y.shape[2:]
    xh = x.permute(0, 2, 3, 1).contiguous(); dyh = dy.permute(0, 2, 3, 1).contiguous()
    taps = [(a, c) for a in range(k[0]) for c in range(k[1])]
    arr = lambda v: (C.c_int32 * len(v))(*v)
    dwk = torch.zeros(k[0], k[1], Cin, Cout, device=dev)
    d = CorrDesc(B=B, H=H, W=W, Cin=Cin, ld_in=Cin, Ho=Ho, Wo=Wo, Cout=Cout, ld_out=Cout, sh=s[0], sw=s[1], out_sw=0, out_sh=0, out_sb=0,
                 n_taps=len(taps), act=0, accumulate=0)
    rc = lib.yad_wgrad_tf32(C.byref(d), arr([a - p[0] for a, _ in taps]), arr([c - p[1] for _, c in taps]), arr([a * k[1] + c for a, c in taps]),
                            xh.data_ptr(), dyh.data_ptr(), dwk.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    print("rc", rc, lib.yad_last_error())
    torch.cuda.synchronize()
    ref = gw.permute(2, 3, 1, 0)
    print((B, H, W, Cin, Cout, k, s, p), "max|ref|", float(ref.abs().max()), "max|got|", float(dwk.abs().max()), "max err", float((dwk - ref).abs().max()))
    if float((dwk - ref).abs().max()) > 0.01 * float(ref.abs().max()):
        # diagnose: which entries match
        r = (dwk - ref).abs() < 0.01 * float(ref.abs().max())
        print("  match frac", float(r.float().mean()), "per tap", r.float().mean(dim=(2, 3)).flatten().tolist())
        print("  match per ci block of 32:", [float(r[:, :, i:i + 32].float().mean()) for i in range(0, Cin, 32)])
        print("  match per co block of 32:", [float(r[..., i:i + 32].float().mean()) for i in range(0, Cout, 32)])
        print("  got[1,1,:4,:4]", dwk[k[0] // 2, k[1] // 2, :4, :4].tolist()); print("  ref[1,1,:4,:4]", ref[k[0] // 2, k[1] // 2, :4, :4].tolist())
run(2, 8, 24, 64, 64, (3, 3), (1, 1), (1, 1))
run(2, 8, 24, 128, 128, (1, 1), (1, 1), (0, 0))
run(2, 8, 24, 32, 32, (1, 1), (1, 1), (0, 0))
run(2, 8, 24, 64, 128, (3, 3), (2, 2), (1, 1))
