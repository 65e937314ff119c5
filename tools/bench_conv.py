#!/usr/bin/env python
"""Micro-benchmark of yad_conv_flat on the backbone's layer shapes (CUDA-event timed, L2 flushed between runs)."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import yad_b200  # noqa: E402,F401
from yad_b200 import _lib  # noqa: E402

SHAPES = {  # name: (H, W, Cin, Cout, k)
    "layer1": (8, 240, 64, 64, 3), "layer2": (4, 120, 128, 128, 3), "layer3": (2, 60, 256, 256, 3), "layer4": (1, 30, 512, 512, 3),
    "l1_1x1": (8, 240, 64, 64, 1), "l2_1x1": (4, 120, 128, 128, 1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--shapes", default="layer1,layer2,layer3,layer4,l1_1x1,l2_1x1")
    ap.add_argument("--timeline", type=int, default=0, help="print the MMA warp's cycle budget (yad_conv_flat_set_timeline)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.init(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in a.shapes.split(","):
        H, W, Cin, Cout, k = SHAPES[name]
        B = a.batch
        Hp, Wp = (H + k // 2 if H > 1 else 1), W + k // 2
        x = torch.randn(B, Wp, Hp, Cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(Cout, k * k * Cin, device=dev) / (Cin * k * k) ** 0.5).to(torch.bfloat16)
        bias = torch.zeros(Cout, device=dev)
        res = torch.randn(B, Wp, Hp, Cout, device=dev).to(torch.bfloat16) if a.res else None
        out = torch.zeros(B, Wp, Hp, Cout, device=dev, dtype=torch.bfloat16)
        d = _lib.FlatDesc(B=B, H=H, W=W, Hp=Hp, Wp=Wp, Cin=Cin, ld_in=Cin, Cout=Cout, ld_out=Cout, co_off=0, kh=k, kw=k, ph=k // 2,
                          pw=k // 2, act=1, ld_res=Cout if a.res else 0)
        ts = []
        for i in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.yad_conv_flat(C.byref(d), x.data_ptr(), w.data_ptr(), Cout, bias.data_ptr(), _lib.ptr(res), out.data_ptr(),
                                         a.flags, st), "conv_flat")
            e1.record()
            torch.cuda.synchronize()
            if i:
                ts.append(e0.elapsed_time(e1) * 1e3)
        if a.timeline:
            tl = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
            _lib.check(lib.yad_conv_flat_set_timeline(tl.data_ptr()), "tl")
            _lib.check(lib.yad_conv_flat(C.byref(d), x.data_ptr(), w.data_ptr(), Cout, bias.data_ptr(), _lib.ptr(res), out.data_ptr(),
                                         a.flags, st), "conv_flat")
            torch.cuda.synchronize()
            _lib.check(lib.yad_conv_flat_set_timeline(None), "tl")
            tv = tl.view(148, 16).double()
            m = tv.mean(0).tolist()
            print(f"   MMA warp cycles (mean over CTAs, max total {tv[:, 0].max().item():.0f}): total {m[0]:.0f}  wait acc {m[1]:.0f}  "
                  f"wait patch {m[2]:.0f}  wait weights {m[3]:.0f}  issue block {m[4]:.0f}")
            print(f"   epilogue thread cycles (mean over CTAs): wait accumulator {m[5]:.0f}  wait staging buffer {m[6]:.0f}  tcgen05.ld {m[7]:.0f}  "
                  f"(epilogue lifetime {(tv[:, 12] - tv[:, 9]).mean():.0f})")
            ent = tv[:, 8:9]
            rel = (tv[:, 9:14] - ent)
            i = int(torch.argmax(rel[:, 4]))
            ns = (tv[:, 15] - tv[:, 14])
            g0 = tv[:, 14].min()
            print(f"   coarse (cycles from CTA entry; mean | slowest CTA {i}): setup {rel[:, 0].mean():.0f} | {rel[i, 0]:.0f}  mma begin {rel[:, 1].mean():.0f} | "
                  f"{rel[i, 1]:.0f}  mma end {rel[:, 2].mean():.0f} | {rel[i, 2]:.0f}  epilogue end {rel[:, 3].mean():.0f} | {rel[i, 3]:.0f}  exit "
                  f"{rel[:, 4].mean():.0f} | {rel[i, 4]:.0f};  CTA lifetime {ns.mean() / 1e3:.1f} us mean, {ns.max() / 1e3:.1f} max; first entry -> last exit "
                  f"{(tv[:, 15].max() - g0) / 1e3:.1f} us; entry spread {(tv[:, 14].max() - g0) / 1e3:.1f} us; clock {rel[i, 4] / ns[i]:.3f} GHz")
        t = sorted(ts)[len(ts) // 2]
        useful = B * H * W * Cout * Cin * (k * k if H > 1 else k) / 1e6
        print(f"{name:8s} B={B} flags={a.flags} res={a.res}: {t:8.1f} us  useful {2 * useful / t / 1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
