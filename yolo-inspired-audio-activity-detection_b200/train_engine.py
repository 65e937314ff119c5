"""Train-mode forward and backward of the network: what autograd does for
``TrainerPipeline.__feed`` (reference: pipeline/_trainer.py:94-108) when the model is in ``train()`` mode -
batch-statistics BatchNorm with running-stat updates, dropout, the train-form RepVGG branches, and the gradient of
every parameter including the three anchor parameters (modules/_architecture.py:39-41).

The graph is the reference's (modules/_backbone.py:142-152, torchvision resnet.py:89-105, modules/_common.py:43-48,
86-95,179-185,204-215,241-265, modules/_architecture.py:132-156); every node is one or two C-ABI calls on fp32 NHWC
tensors and records a closure on a tape, the backward replays the tape in reverse.  torch supplies device memory, the
stream and the autograd hook (`_TrainFn`), nothing else.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, F32, ConvDesc, CorrDesc


def _c16(v: int) -> int:
    return (v + 15) // 16 * 16


def _c32(v: int) -> int:
    return (v + 31) // 32 * 32


class _T:
    """A channel slice [off, off + C) of an fp32 NHWC buffer [B, H, W, ld]."""
    __slots__ = ("buf", "off", "C")

    def __init__(self, buf: torch.Tensor, off: int = 0, C: Optional[int] = None):
        self.buf, self.off = buf, off
        self.C = buf.shape[3] - off if C is None else C

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + 4 * self.off

    @property
    def ld(self) -> int:
        return self.buf.shape[3]

    @property
    def B(self):
        return self.buf.shape[0]

    @property
    def H(self):
        return self.buf.shape[1]

    @property
    def W(self):
        return self.buf.shape[2]

    @property
    def rows(self) -> int:
        return self.buf.shape[0] * self.buf.shape[1] * self.buf.shape[2]


class TrainEngine:
    def __init__(self, model, device: torch.device, conv_mode: str = "tf32"):
        if conv_mode not in ("tf32", "f32"):
            raise ValueError("train conv_mode must be 'tf32' (tcgen05 tensor cores, cuDNN's default for fp32 training) or 'f32'")
        self.conv_mode = conv_mode
        if device.type != "cuda":
            raise RuntimeError("yad_b200 train mode needs the model on a CUDA (sm_100a) device; there is no CPU fallback")
        self.dev = device
        self.lib = _lib.init(device.index if device.index is not None else torch.cuda.current_device())
        self.model = model
        self.cfg = model.config
        self.nc = model.num_classes
        self.A = self.cfg["num_anchors"]
        self.E = 3 + self.nc
        self._wcache: Dict[object, tuple] = {}
        self._plans: Dict[tuple, dict] = {}
        self._step = 0
        self._seed_dev = torch.zeros(1, device=device, dtype=torch.int64)     # train-step counter (dropout seed offset)
        self._graphs: Dict[tuple, dict] = {}
        self._force_repack = False
        self._bulk_packed = False          # the packed TF32 weights were refreshed by _bulk_pack() for the forward in progress
        self._pack_entries: Dict = {}
        self._pack_table = None
        self._bwd_arena, self._bwd_used, self._stats_arena, self._stats_used, self._act_numel = None, 0, None, 0, 0
        self._n_weights = sum(p.numel() for p in model.parameters())
        self._n_bn = sum(m.num_features for m in model.modules() if isinstance(m, nn.BatchNorm2d))
        # split-K (fp32 red.add of partial sums) for small forward grids: faster, but the summation order then varies from run
        # to run and TF32 operand truncation amplifies that 1e-7 noise to ~1e-3 after 20 layers; off = reproducible forward
        self.fwd_split_k = os.environ.get("YAD_TRAIN_FWD_SPLITK", "0") == "1"
        for m in model.modules():
            if hasattr(m, "conv_reparam"):
                raise NotImplementedError("yad_b200: train-mode forward of the deploy (re-parameterised) form is not built")

    # ------------------------------------------------------------------ helpers
    def _s(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _zeros(self, numel: int) -> torch.Tensor:
        """fp32 zeros carved from the backward arena (one memset per step instead of one per gradient buffer)."""
        n = (numel + 63) // 64 * 64
        ar = self._bwd_arena
        if ar is None or self._bwd_used + n > ar.numel():
            return torch.zeros(numel, device=self.dev, dtype=torch.float32)
        t = ar[self._bwd_used:self._bwd_used + numel]
        self._bwd_used += n
        return t

    def _stats(self, Cc: int) -> torch.Tensor:
        """[2*C] fp64 zeros for the BatchNorm moments, carved from the forward's statistics arena."""
        ar = self._stats_arena
        if ar is None or self._stats_used + 2 * Cc > ar.numel():
            return torch.zeros(2 * Cc, device=self.dev, dtype=torch.float64)
        t = ar[self._stats_used:self._stats_used + 2 * Cc]
        self._stats_used += 2 * Cc
        return t

    def _new(self, B, H, W, Cc, zero=False) -> _T:
        """Fresh activation buffer.  Channel counts that are not a multiple of 32 (the 15-channel head tensors) get a
        zero-filled pitch of 32 so that the TF32 kernels can read whole 128-byte rows."""
        ld = Cc if Cc % 32 == 0 or Cc < 8 else (Cc + 31) // 32 * 32
        f = torch.zeros if (zero or ld != Cc) else torch.empty
        self._act_numel += (B * H * W * ld + 63) // 64 * 64
        return _T(f((B, H, W, ld), device=self.dev, dtype=torch.float32), 0, Cc)

    def _grad(self, t: _T) -> _T:
        """Gradient slice matching ``t`` (one zero-initialised buffer per activation buffer; every backward accumulates)."""
        key = t.buf.data_ptr()
        g = self._grads.get(key)
        if g is None:
            g = self._grads[key] = self._zeros(t.buf.numel()).view(t.buf.shape)
        return _T(g, t.off, t.C)

    @staticmethod
    def _pgrad(p: torch.Tensor) -> torch.Tensor:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        return p.grad

    def _permute(self, src_ptr: int, strides, out: torch.Tensor, sizes, accumulate: bool):
        st = (C.c_int64 * 4)(*strides)
        sz = (C.c_int32 * 4)(*sizes)
        _lib.check(self.lib.yad_permute4(src_ptr, st, out.data_ptr(), sz, 1 if accumulate else 0, self._s()), "permute4")

    def _packed(self, w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """OIHW master weight -> ([kh][kw][Cin][Cout] for the forward / wgrad, [kh][kw][Cout][Cin] for the dgrad)."""
        ent = self._wcache.get(id(w))
        ver = (w._version, _lib.param_epoch, w.data_ptr())
        if ent is not None and ent[0] == ver:
            return ent[1], ent[2]
        O, I, kh, kw = w.shape
        wf = torch.empty((kh, kw, I, O), device=self.dev, dtype=torch.float32)
        wt = torch.empty((kh, kw, O, I), device=self.dev, dtype=torch.float32)
        wd = w.detach()
        if not wd.is_contiguous():
            wd = wd.contiguous()
        self._permute(wd.data_ptr(), (kw, 1, kh * kw, I * kh * kw), wf, (kh, kw, I, O), False)
        self._permute(wd.data_ptr(), (kw, 1, I * kh * kw, kh * kw), wt, (kh, kw, O, I), False)
        self._wcache[id(w)] = (ver, wf, wt)
        return wf, wt

    # ------------------------------------------------------------------ nodes (forward + recorded backward)
    def conv(self, x: _T, conv: nn.Conv2d, out: Optional[_T] = None, need_dx: bool = True) -> _T:
        I = conv.weight.shape[1]
        if self.conv_mode == "tf32" and I >= 8 and x.off == 0 and x.ld % 32 == 0 and (I % 32 == 0 or x.ld == _c32(I)):
            return self._conv_tf32(x, conv, out, need_dx)
        return self._conv_f32(x, conv, out, need_dx)

    def _conv_geom(self, x: _T, conv: nn.Conv2d, out: Optional[_T]):
        O, I, kh, kw = conv.weight.shape
        sh, sw = conv.stride
        ph, pw = conv.padding
        assert x.C == I, (x.C, I)
        Ho, Wo = (x.H + 2 * ph - kh) // sh + 1, (x.W + 2 * pw - kw) // sw + 1
        if out is None:
            out = self._new(x.B, Ho, Wo, O)
        assert (out.B, out.H, out.W, out.C) == (x.B, Ho, Wo, O)
        return O, I, kh, kw, sh, sw, ph, pw, Ho, Wo, out

    def _wgrad_finish(self, w, b, dwk, dy: _T, O, I, kh, kw):
        """dwk [kh][kw][Cin][Cout] -> += OIHW .grad; bias gradient = column sums of dY."""
        self._permute(dwk.data_ptr(), (1, O, kw * I * O, I * O), self._pgrad(w), (O, I, kh, kw), True)
        if b is not None:
            bws = torch.zeros(O, device=self.dev, dtype=torch.float64)
            _lib.check(self.lib.yad_colsum_f64(dy.ptr, dy.ld, dy.rows, O, bws.data_ptr(), self._s()), "bias grad (column sums)")
            _lib.check(self.lib.yad_add_f64_to_f32(bws.data_ptr(), O, self._pgrad(b).data_ptr(), self._s()), "bias grad")

    def _conv_f32(self, x: _T, conv: nn.Conv2d, out: Optional[_T], need_dx: bool) -> _T:
        """fp32 CUDA-core path (parity mode; always used for the 2-channel stem conv1)."""
        w, b = conv.weight, conv.bias
        O, I, kh, kw, sh, sw, ph, pw, Ho, Wo, out = self._conv_geom(x, conv, out)
        wf, wt = self._packed(w)
        d = ConvDesc(B=x.B, H=x.H, W=x.W, Cin=I, ld_in=x.ld, Cout=O, ld_out=out.ld, co_off=0, kh=kh, kw=kw, sh=sh, sw=sw,
                     ph=ph, pw=pw, act=ACT_NONE, ld_res=0)
        _lib.check(self.lib.yad_conv_simt(C.byref(d), F32, x.ptr, wf.data_ptr(), O, _lib.ptr(b), 0, out.ptr, self._s()), "conv fwd")
        if self._tape is not None:
            def bwd():
                dy = self._grad(out)
                dwk = torch.zeros_like(wf)
                _lib.check(self.lib.yad_conv_wgrad(C.byref(d), x.ptr, dy.ptr, dwk.data_ptr(), 0, self._s()), "conv wgrad")
                self._wgrad_finish(w, b, dwk, dy, O, I, kh, kw)
                if need_dx:
                    dx = self._grad(x)
                    _lib.check(self.lib.yad_conv_dgrad(C.byref(d), dy.ptr, wt.data_ptr(), dx.ptr, dx.ptr, self._s()), "conv dgrad")
            self._tape.append(bwd)
        return out

    def _packed_tf32(self, w: torch.Tensor, cin_pad: int, coutk: int):
        """OIHW master weight -> K-major fp32 operands of the TF32 kernels: forward [ceil16(O)][kh*kw*cin_pad] and data
        gradient [ceil16(I)][kh*kw*coutk] (un-flipped taps; the tap list carries the geometry)."""
        key = ("tf32", id(w))
        ver = (w._version, _lib.param_epoch, w.data_ptr(), cin_pad, coutk)
        ent = self._wcache.get(key)
        if ent is not None and (self._bulk_packed or (ent[0] == ver and not self._force_repack)):
            return ent[1], ent[2]          # up to date (or refreshed by the one-launch bulk pack at the start of this forward)
        O, I, kh, kw = w.shape
        wd = w.detach()
        if ent is not None and ent[1].shape == (_c16(O), kh, kw, cin_pad) and ent[2].shape == (_c16(I), kh, kw, coutk):
            wf, wt = ent[1], ent[2]          # pad entries are already zero
        else:
            wf = torch.zeros((_c16(O), kh, kw, cin_pad), device=self.dev, dtype=torch.float32)
            wt = torch.zeros((_c16(I), kh, kw, coutk), device=self.dev, dtype=torch.float32)
        wf[:O, :, :, :I].copy_(wd.permute(0, 2, 3, 1))
        wt[:I, :, :, :O].copy_(wd.permute(1, 2, 3, 0))
        self._wcache[key] = (ver, wf, wt)
        self._pack_entries[key] = (w, wf, wt, cin_pad, coutk)
        self._pack_table = None
        return wf, wt

    def _ensure_pack_table(self):
        """Device table of yad_pack_weights_tf32 (pointers + geometry of every convolution weight); call outside graph capture."""
        import numpy as np
        if not self._pack_entries:
            return
        ptrs = [w.data_ptr() for w, *_ in self._pack_entries.values()]
        if self._pack_table is not None and self._pack_table[3] == ptrs:
            return
        dt = np.dtype([("src", "<u8"), ("wf", "<u8"), ("wt", "<u8"), ("O", "<i4"), ("I", "<i4"), ("KK", "<i4"), ("cin_pad", "<i4"),
                       ("coutk", "<i4"), ("pad", "<i4"), ("start", "<i8")])
        assert dt.itemsize == 56
        tab = np.zeros(len(self._pack_entries), dt)
        start = 0
        for j, (w, wf, wt, cin_pad, coutk) in enumerate(self._pack_entries.values()):
            O, I, kh, kw = w.shape
            if not w.is_contiguous() or wf.shape[3] != cin_pad or wt.shape[3] != coutk:
                self._pack_table = None
                return
            tab[j] = (w.data_ptr(), wf.data_ptr(), wt.data_ptr(), O, I, kh * kw, cin_pad, coutk, 0, start)
            start += O * I * kh * kw
        self._pack_table = (torch.from_numpy(tab.view(np.uint8).copy()).to(self.dev), len(tab), start, ptrs)

    def _bulk_pack(self):
        """Refresh every packed TF32 weight operand with ONE launch (yad_pack_weights_tf32) instead of two strided copies per
        convolution; used when the step is captured into CUDA graphs (the parameters change every step, so the packing is part
        of the graph).  The table holds device pointers: parameters and packed buffers must stay put, which the graphs need anyway."""
        if self._pack_table is None:       # built outside graph capture (it needs a host -> device copy): _ensure_pack_table
            return False
        t, n, total, ptrs = self._pack_table
        if ptrs != [w.data_ptr() for w, *_ in self._pack_entries.values()]:
            return False
        _lib.check(self.lib.yad_pack_weights_tf32(t.data_ptr(), n, total, self._s()), "pack_weights_tf32")
        return True

    def _conv_tf32(self, x: _T, conv: nn.Conv2d, out: Optional[_T], need_dx: bool, stats: Optional[torch.Tensor] = None) -> _T:
        """tcgen05 kind::tf32 path: forward, data gradient (one correlation per output stride-parity class) and weight
        gradient (MN-major operands) - see csrc/conv_tf32.cu."""
        w, b = conv.weight, conv.bias
        O, I, kh, kw, sh, sw, ph, pw, Ho, Wo, out = self._conv_geom(x, conv, out)
        cin_pad = _c32(I)
        plan = self._plans.get((id(conv), x.B, x.H, x.W, x.ld, out.ld))
        if plan is None:
            arr = lambda v: (C.c_int32 * len(v))(*v)      # noqa: E731
            taps = [(a, c) for a in range(kh) for c in range(kw)]
            plan = {"fwd": (arr([a - ph for a, _ in taps]), arr([c - pw for _, c in taps]), arr([(a * kw + c) * cin_pad for a, c in taps])),
                    "dst": arr([a * kw + c for a, c in taps]), "dgrad": []}
            coutk = _c32(O)
            for rh in range(sh):
                for rw in range(sw):
                    Hc, Wc = (x.H - rh + sh - 1) // sh, (x.W - rw + sw - 1) // sw
                    tp = [(a, c) for a, c in taps if (rh + ph - a) % sh == 0 and (rw + pw - c) % sw == 0]
                    if Hc <= 0 or Wc <= 0:
                        continue
                    plan["dgrad"].append((rh, rw, Hc, Wc, len(tp), arr([(rh + ph - a) // sh for a, _ in tp]),
                                          arr([(rw + pw - c) // sw for _, c in tp]), arr([(a * kw + c) * coutk for a, c in tp])))
            self._plans[(id(conv), x.B, x.H, x.W, x.ld, out.ld)] = plan
        assert out.off % 4 == 0
        coutk = _c32(O)
        wf, wt = self._packed_tf32(w, cin_pad, coutk)
        d = CorrDesc(B=x.B, H=x.H, W=x.W, Cin=cin_pad, ld_in=x.ld, Ho=Ho, Wo=Wo, Cout=O, ld_out=out.ld, sh=sh, sw=sw, out_sw=0,
                     out_sh=0, out_sb=0, n_taps=kh * kw, act=ACT_NONE, accumulate=0,
                     whole_rows=1 if (self.fwd_split_k and out.off == 0 and out.ld in (O, _c32(O))) else 0)
        fdh, fdw, fk = plan["fwd"]
        _lib.check(self.lib.yad_corr_tf32(C.byref(d), fdh, fdw, fk, x.ptr, wf.data_ptr(), wf.shape[0], kh * kw * cin_pad, _lib.ptr(b),
                                          out.ptr, _lib.ptr(stats), self._s()), "conv fwd (tf32)")
        if self._tape is not None:
            def bwd():
                dy = self._grad(out)
                assert dy.off == 0 and dy.ld == coutk, "conv outputs are whole buffers with a 32-channel pitch"
                dwk = self._zeros(kh * kw * I * O)
                dg = CorrDesc(B=x.B, H=x.H, W=x.W, Cin=I, ld_in=x.ld, Ho=Ho, Wo=Wo, Cout=O, ld_out=dy.ld, sh=sh, sw=sw, out_sw=0,
                              out_sh=0, out_sb=0, n_taps=kh * kw, act=ACT_NONE, accumulate=0)
                _lib.check(self.lib.yad_wgrad_tf32(C.byref(dg), fdh, fdw, plan["dst"], x.ptr, dy.ptr, dwk.data_ptr(), self._s()), "conv wgrad (tf32)")
                self._wgrad_finish(w, b, dwk, dy, O, I, kh, kw)
                if need_dx:
                    dx = self._grad(x)
                    for rh, rw, Hc, Wc, nt, tdh, tdw, tk in plan["dgrad"]:
                        if nt == 0:
                            continue      # no tap reaches this parity class: its gradient is zero
                        dd = CorrDesc(B=x.B, H=Ho, W=Wo, Cin=coutk, ld_in=dy.ld, Ho=Hc, Wo=Wc, Cout=I, ld_out=dx.ld, sh=1, sw=1,
                                      out_sw=sw, out_sh=sh * x.W, out_sb=x.H * x.W, n_taps=nt, act=ACT_NONE, accumulate=1)
                        _lib.check(self.lib.yad_corr_tf32(C.byref(dd), tdh, tdw, tk, dy.ptr, wt.data_ptr(), wt.shape[0], kh * kw * coutk,
                                                          0, dx.ptr + 4 * (rh * x.W + rw) * dx.ld, 0, self._s()), "conv dgrad (tf32)")
            self._tape.append(bwd)
        return out

    def stem_tf32(self, xs: torch.Tensor, conv: nn.Conv2d) -> _T:
        """conv1 (2 -> 64, 7x7, stride 2, no bias; modules/_backbone.py:143) on the tensor cores: im2col to K = 128 (98 real
        taps x channels) and a 1x1 TF32 conv; its weight gradient is the 1x1 TF32 wgrad on the same patches."""
        w = conv.weight
        O, I, kh, kw = w.shape
        assert (kh, kw) == (7, 7) and tuple(conv.stride) == (2, 2) and tuple(conv.padding) == (3, 3) and conv.bias is None
        B, _, H, W = xs.shape
        K = _c32(kh * kw * I)
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        patches = _T(torch.empty((B, Ho, Wo, K), device=self.dev, dtype=torch.float32))
        _lib.check(self.lib.yad_stem_im2col(xs.data_ptr(), B, I, H, W, K, patches.ptr, self._s()), "stem im2col")
        key = ("stem", id(w))
        ver = (w._version, _lib.param_epoch, w.data_ptr())
        ent = self._wcache.get(key)
        if ent is None or ent[0] != ver or self._force_repack:
            wf = ent[1] if ent is not None else torch.zeros((_c16(O), K), device=self.dev, dtype=torch.float32)
            wf[:O, :kh * kw * I].copy_(w.detach().permute(0, 2, 3, 1).reshape(O, -1))
            self._wcache[key] = ent = (ver, wf)
        wf = ent[1]
        out = self._new(B, Ho, Wo, O)
        z = (C.c_int32 * 1)(0)
        d = CorrDesc(B=B, H=Ho, W=Wo, Cin=K, ld_in=K, Ho=Ho, Wo=Wo, Cout=O, ld_out=out.ld, sh=1, sw=1, out_sw=0, out_sh=0, out_sb=0,
                     n_taps=1, act=ACT_NONE, accumulate=0, whole_rows=0)
        _lib.check(self.lib.yad_corr_tf32(C.byref(d), z, z, z, patches.ptr, wf.data_ptr(), wf.shape[0], K, 0, out.ptr, 0, self._s()), "stem conv1 (tf32)")
        if self._tape is not None:
            def bwd():
                dy = self._grad(out)
                dwk = self._zeros(K * O)
                _lib.check(self.lib.yad_wgrad_tf32(C.byref(d), z, z, z, patches.ptr, dy.ptr, dwk.data_ptr(), self._s()), "stem wgrad (tf32)")
                self._wgrad_finish(w, None, dwk, dy, O, I, kh, kw)       # rows 0..97 of dwk are [kh][kw][Cin][Cout]
            self._tape.append(bwd)
        return out

    def bn(self, x: _T, bn: nn.BatchNorm2d, act: int, out: Optional[_T] = None, sums: Optional[torch.Tensor] = None) -> _T:
        Cc, N = x.C, x.rows
        if out is None:
            out = self._new(x.B, x.H, x.W, Cc)
        sm = torch.empty(2 * Cc, device=self.dev, dtype=torch.float32)
        mean, invstd = sm[:Cc], sm[Cc:]
        mom = 0.1 if bn.momentum is None else float(bn.momentum)
        if sums is not None:
            ws = sums            # the backward reuses it as scratch
            _lib.check(self.lib.yad_bn_train_apply(x.ptr, x.ld, N, Cc, bn.weight.data_ptr(), bn.bias.data_ptr(), float(bn.eps), mom,
                                                   bn.running_mean.data_ptr(), bn.running_var.data_ptr(), act, out.ptr, out.ld,
                                                   mean.data_ptr(), invstd.data_ptr(), sums.data_ptr(), self._s()), "bn apply")
        else:
            ws = torch.empty(2 * Cc, device=self.dev, dtype=torch.float64)
            _lib.check(self.lib.yad_bn_train_fwd(x.ptr, x.ld, N, Cc, bn.weight.data_ptr(), bn.bias.data_ptr(), float(bn.eps), mom,
                                                 bn.running_mean.data_ptr(), bn.running_var.data_ptr(), act, out.ptr, out.ld,
                                                 mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(), self._s()), "bn fwd")
        self._bn_counters.append(bn.num_batches_tracked)
        if self._tape is not None:
            def bwd():
                dy, dx = self._grad(out), self._grad(x)
                _lib.check(self.lib.yad_bn_train_bwd(x.ptr, x.ld, out.ptr, out.ld, dy.ptr, dy.ld, N, Cc, bn.weight.data_ptr(),
                                                     mean.data_ptr(), invstd.data_ptr(), act, dx.ptr, dx.ld, 1,
                                                     self._pgrad(bn.weight).data_ptr(), self._pgrad(bn.bias).data_ptr(),
                                                     ws.data_ptr(), self._s()), "bn bwd")
            self._tape.append(bwd)
        return out

    def conv_bn(self, x: _T, conv: nn.Conv2d, bn: nn.BatchNorm2d, act: int, out: Optional[_T] = None) -> _T:
        """conv -> BatchNorm(batch statistics) -> activation.  On the TF32 path the moments come out of the conv epilogue, so
        the BatchNorm forward is one launch (finalise + running statistics + apply)."""
        I = conv.weight.shape[1]
        if not (self.conv_mode == "tf32" and I >= 8 and x.off == 0 and x.ld % 32 == 0 and (I % 32 == 0 or x.ld == _c32(I))):
            return self.bn(self.conv(x, conv), bn, act, out)
        sums = self._stats(conv.weight.shape[0])
        y = self._conv_tf32(x, conv, None, True, stats=sums)
        return self.bn(y, bn, act, out, sums=sums)

    def cbl(self, x: _T, m, out: Optional[_T] = None) -> _T:
        """ConvBorINorm (modules/_common.py:43-48): conv(+bias) -> BatchNorm (batch statistics) -> LeakyReLU(0.2)."""
        return self.conv_bn(x, m.conv, m.norm, ACT_LRELU if m.has_activation else ACT_NONE, out)

    def add_act(self, a: _T, b: _T, c: Optional[_T], act: int, out: Optional[_T] = None) -> _T:
        if out is None:
            out = self._new(a.B, a.H, a.W, a.C)
        N, Cc = a.rows, a.C
        _lib.check(self.lib.yad_add_act(a.ptr, a.ld, b.ptr, b.ld, 0 if c is None else c.ptr, 0 if c is None else c.ld, N, Cc, act,
                                        out.ptr, out.ld, self._s()), "add_act")
        if self._tape is not None:
            def bwd():
                dy = self._grad(out)
                da, db = self._grad(a), self._grad(b)
                dc = self._grad(c) if c is not None else None
                _lib.check(self.lib.yad_add_act_bwd(out.ptr, out.ld, dy.ptr, dy.ld, N, Cc, act, da.ptr, da.ld, db.ptr, db.ld,
                                                    0 if dc is None else dc.ptr, 0 if dc is None else dc.ld, self._s()), "add_act bwd")
            self._tape.append(bwd)
        return out

    def dropout(self, x: _T, p: float) -> _T:
        """nn.Dropout (modules/_backbone.py:133,147) with a counter-based mask; the step counter lives on the device so that
        a captured CUDA graph draws a fresh mask on every replay."""
        if p <= 0.0:
            return x
        assert x.off == 0 and x.C == x.ld
        out = self._new(x.B, x.H, x.W, x.C)
        seed = (int(torch.initial_seed()) * 1000003) & ((1 << 62) - 1)
        n = x.buf.numel()
        sd = self._seed_snap
        _lib.check(self.lib.yad_dropout_dev(x.ptr, n, p, seed, sd.data_ptr(), 0, out.ptr, self._s()), "dropout")
        if self._tape is not None:
            def bwd():
                _lib.check(self.lib.yad_dropout_dev(self._grad(out).ptr, n, p, seed, sd.data_ptr(), 1, self._grad(x).ptr, self._s()), "dropout bwd")
            self._tape.append(bwd)
        return out

    def hmean(self, x: _T) -> _T:
        """adaptive_avg_pool2d(H -> 1), modules/_common.py:248-252."""
        if x.H == 1:
            return x
        assert x.off == 0 and x.C == x.ld
        B, H, W, Cc = x.B, x.H, x.W, x.C
        out = self._new(B, 1, W, Cc)
        _lib.check(self.lib.yad_hmean(x.ptr, F32, B, H, W, Cc, Cc, 1, W, H * W, out.ptr, Cc, 0, self._s()), "hmean")
        if self._tape is not None:
            def bwd():
                _lib.check(self.lib.yad_hmean_bwd(self._grad(out).ptr, Cc, B, H, W, Cc, self._grad(x).ptr, Cc, self._s()), "hmean bwd")
            self._tape.append(bwd)
        return out

    def resize_w(self, x: _T, up: bool, out: _T) -> _T:
        """F.interpolate(scale_factor=(1, 2 | 0.5), bilinear), modules/_common.py:181-182."""
        B, W, Cc = x.B, x.W, x.C
        assert out.C == Cc and out.W == (2 * W if up else W // 2)
        _lib.check(self.lib.yad_resize_w(x.buf.data_ptr(), F32, B, W, Cc, x.ld, x.off, 1 if up else 0, out.buf.data_ptr(), out.ld,
                                         out.off, self._s()), "resize_w")
        if self._tape is not None:
            def bwd():
                dy, dx = self._grad(out), self._grad(x)
                _lib.check(self.lib.yad_resize_w_bwd(dy.ptr, dy.ld, B, W, Cc, 1 if up else 0, dx.ptr, dx.ld, self._s()), "resize_w bwd")
            self._tape.append(bwd)
        return out

    def maxpool5(self, x: _T, out: _T) -> _T:
        B, W, Cc = x.B, x.W, x.C
        _lib.check(self.lib.yad_maxpool5_w(x.ptr, x.ld, B, W, Cc, out.ptr, out.ld, self._s()), "maxpool5")
        if self._tape is not None:
            def bwd():
                dy, dx = self._grad(out), self._grad(x)
                _lib.check(self.lib.yad_maxpool5_w_bwd(x.ptr, x.ld, dy.ptr, dy.ld, B, W, Cc, dx.ptr, dx.ld, self._s()), "maxpool5 bwd")
            self._tape.append(bwd)
        return out

    def repvgg(self, x: _T, m, out: Optional[_T] = None) -> _T:
        """Train-form RepVGG block (modules/_common.py:86-95): lrelu(cbl3x3(x) + cbl1x1(x) [+ bn(x)])."""
        a = self.cbl(x, m.conv3x3)
        b = self.cbl(x, m.conv1x1)
        c = self.bn(x, m.identity, ACT_NONE) if isinstance(m.identity, nn.BatchNorm2d) else None
        return self.add_act(a, b, c, ACT_LRELU, out)

    def repblock(self, x: _T, rb, out: Optional[_T] = None) -> _T:
        blocks = [rb.conv1] + (list(rb.blocks) if isinstance(rb.blocks, nn.Sequential) else [])
        for i, blk in enumerate(blocks):
            x = self.repvgg(x, blk, out if i == len(blocks) - 1 else None)
        return x

    def basic_block(self, x: _T, blk) -> _T:
        t = self.conv_bn(x, blk.conv1, blk.bn1, ACT_RELU)
        u = self.conv_bn(t, blk.conv2, blk.bn2, ACT_NONE)
        idt = x if blk.downsample is None else self.conv_bn(x, blk.downsample[0], blk.downsample[1], ACT_NONE)
        return self.add_act(u, idt, None, ACT_RELU)

    def bottleneck(self, x: _T, blk) -> _T:
        """torchvision Bottleneck in train mode (resnet.py:143-163 via modules/_backbone.py:128-138, ``resnet_config.block:
        Bottleneck``): 1x1 - 3x3 (carries the stride) - 1x1 (x4 channels), each with batch-statistics BatchNorm; ReLU after the
        residual add."""
        t = self.conv_bn(x, blk.conv1, blk.bn1, ACT_RELU)
        u = self.conv_bn(t, blk.conv2, blk.bn2, ACT_RELU)
        v = self.conv_bn(u, blk.conv3, blk.bn3, ACT_NONE)
        idt = x if blk.downsample is None else self.conv_bn(x, blk.downsample[0], blk.downsample[1], ACT_NONE)
        return self.add_act(v, idt, None, ACT_RELU)

    # ------------------------------------------------------------------ the graph
    def forward(self, xs: torch.Tensor, T: int, L_res: int, record: bool):
        """x_spectral [B,2,32,T] f32 (frontend output, not differentiable: it has no parameters) -> (preds per scale
        [B,G,A,E], tape)."""
        model = self.model
        fe, ms = model.feature_extractor, model.multiscale_module
        # graph capture: refresh all packed TF32 weights with one launch; the per-convolution copies are then skipped
        self._bulk_packed = bool(self._force_repack and self.conv_mode == "tf32" and self._pack_entries
                                 and os.environ.get("YAD_BULK_PACK", "1") != "0" and self._bulk_pack())
        self._tape: Optional[List[Callable[[], None]]] = [] if record else None
        self._grads: Dict[int, torch.Tensor] = {}
        self._bn_counters: List[torch.Tensor] = []
        self._step += 1
        self._seed_dev.add_(1)
        # per-forward snapshot of the counter: the backward of THIS tape must redraw the mask its forward used, even when another
        # train-mode forward (gradient accumulation, a no_grad train-mode pass) advances the counter in between.  Inside a graph
        # capture the clone is a copy node into a pool tensor, replayed with the graph.
        self._seed_snap = self._seed_dev.clone()
        self._act_numel = 0
        self._stats_arena, self._stats_used = torch.zeros(2 * self._n_bn, device=self.dev, dtype=torch.float64), 0
        B, Cin, H0, _ = xs.shape
        if self.conv_mode == "tf32":
            x = self.stem_tf32(xs, fe.conv1)
        else:
            x0 = self._new(B, H0, T, Cin)
            self._permute(xs.data_ptr(), (Cin * H0 * T, T, 1, H0 * T), x0.buf, (B, H0, T, Cin), False)     # NCHW -> NHWC
            x = self.conv(x0, fe.conv1, need_dx=False)
        x = self.conv_bn(x, fe.conv2, fe.bn1, ACT_RELU)
        x = self.dropout(x, float(fe.dropout_p))
        fmaps, marks = [], []
        for li in range(1, 5):
            for blk in getattr(fe, f"layer{li}"):
                x = self.bottleneck(x, blk) if hasattr(blk, "conv3") else self.basic_block(x, blk)
            fmaps.append(x)
            marks.append(len(self._tape) if self._tape is not None else 0)      # tape length after layer li (gradient buckets)
        hs = [f.H for f in fmaps]
        if not (hs[0] != hs[1] != hs[2] != hs[3]):
            raise NotImplementedError("neck with equal feature-map heights (2-D neck) is not built")
        f1, f2, f3, f4 = [self.hmean(f) for f in fmaps]
        W1, W2, W3, W4 = f1.W, f2.W, f3.W, f4.W
        nh = self.A * self.E
        # CSPSPPF (modules/_common.py:204-215); torch.cat sites are channel slices of one buffer
        sp = ms.cspsppf
        ch = sp.conv2.conv.out_channels
        cat5 = self._new(B, 1, W4, 4 * ch)
        cat7 = self._new(B, 1, W4, 2 * ch)
        cat_n4 = self._new(B, 1, W4, 2 * 128)
        x1 = self.cbl(self.cbl(self.cbl(f4, sp.conv_1_3_4[0]), sp.conv_1_3_4[1]), sp.conv_1_3_4[2], _T(cat5.buf, 0, ch))
        self.cbl(f4, sp.conv2, _T(cat7.buf, ch, ch))
        p1 = self.maxpool5(x1, _T(cat5.buf, ch, ch))
        p2 = self.maxpool5(p1, _T(cat5.buf, 2 * ch, ch))
        self.maxpool5(p2, _T(cat5.buf, 3 * ch, ch))
        self.cbl(self.cbl(cat5, sp.conv5), sp.conv6, _T(cat7.buf, 0, ch))
        p4 = self.cbl(cat7, sp.conv7, _T(cat_n4.buf, 0, 128))
        # BiC3 -> RepBlock3_1 -> p3
        cat_b3 = self._new(B, 1, W3, 256)
        cat_n3 = self._new(B, 1, W3, 256)
        self.cbl(f3, ms.bic3.conv_c1, _T(cat_b3.buf, 0, 64))
        self.resize_w(self.cbl(f2, ms.bic3.conv_c0), False, _T(cat_b3.buf, 64, 64))
        self.resize_w(p4, True, _T(cat_b3.buf, 128, 128))
        p3 = self.repblock(self.cbl(cat_b3, ms.bic3.conv_out), ms.rep_block3_1, _T(cat_n3.buf, 0, 128))
        # BiC2 -> RepBlock2_1 -> n2
        cat_b2 = self._new(B, 1, W2, 256)
        self.cbl(f2, ms.bic2.conv_c1, _T(cat_b2.buf, 0, 64))
        self.resize_w(self.cbl(f1, ms.bic2.conv_c0), False, _T(cat_b2.buf, 64, 64))
        self.resize_w(p3, True, _T(cat_b2.buf, 128, 128))
        n2 = self.repblock(self.cbl(cat_b2, ms.bic2.conv_out), ms.rep_block2_1)
        self.cbl(n2, ms.conv2_downsample, _T(cat_n3.buf, 128, 128))
        n3 = self.repblock(cat_n3, ms.rep_block3_2)
        self.cbl(n3, ms.conv3_downsample, _T(cat_n4.buf, 128, 128))
        n4 = self.repblock(cat_n4, ms.rep_block4_1)
        if self._bn_counters:
            torch._foreach_add_(self._bn_counters, 1)
            _lib.param_epoch += 1      # running statistics were rewritten by the kernels: packed eval engines are stale
        # anchor decode (modules/_architecture.py:132-156), one call per scale
        dur = float(self.cfg["sample_duration"])
        anchors = [model.sm_anchors, model.md_anchors, model.lg_anchors]
        anc_s = torch.stack([a.detach() for a in anchors]).float() * dur          # [3, A] seconds, stays on the device
        center_scaler = T / (L_res / self.cfg["new_sample_rate"])
        preds, heads = [], (n2, n3, n4)
        for s, h in enumerate(heads):
            G = h.W
            assert h.C == nh
            p = torch.empty((B, G, self.A, self.E), device=self.dev, dtype=torch.float32)
            hp = (C.c_void_p * 1)(h.ptr)
            _lib.check(self.lib.yad_decode_dev(hp, (C.c_int32 * 1)(G), (C.c_int32 * 1)(h.ld), (C.c_int32 * 1)(T // G), 1, F32,
                                               anc_s[s].data_ptr(), self.A, self.nc, center_scaler, dur, B, p.data_ptr(), self._s()), "decode")
            preds.append(p)
        state = None
        if record:
            state = {"tape": self._tape, "grads": self._grads, "heads": heads, "anc_s": anc_s, "anchors": anchors,
                     "strides": [T // h.W for h in heads], "center_scaler": center_scaler, "dur": dur, "B": B,
                     "arena_numel": self._act_numel + self._n_weights + 4096 * 64, "marks": marks}
        self._tape, self._grads = None, {}
        self._bulk_packed = False
        return preds, state

    # Gradient buckets, in the order the backward completes them (reverse of the forward): 0 = neck + layer4 (36.6 MB of the
    # 48.5 MB fp32 gradient, finished after ~15 % of the backward's time), 1 = layer3 (8.4 MB), 2 = layer2, layer1, stem and the
    # anchors.  ``on_bucket(i)`` (set by FusedAdamEMA.overlap_allreduce) is called right after bucket i's last kernel was
    # enqueued, so its NCCL all-reduce runs under the rest of the backward ("bucketed behind backward").
    N_BUCKETS = 3
    on_bucket: Optional[Callable[[int], None]] = None

    def bucket_params(self) -> List[List[torch.nn.Parameter]]:
        fe, ms = self.model.feature_extractor, self.model.multiscale_module
        b0 = list(fe.layer4.parameters()) + list(ms.parameters())
        b1 = list(fe.layer3.parameters())
        seen = {id(p) for p in b0 + b1}
        b2 = [p for p in self.model.parameters() if id(p) not in seen]
        return [b0, b1, b2]

    def backward_segments(self, state, dpreds: List[Optional[torch.Tensor]]) -> List[Callable[[], None]]:
        """The backward as N_BUCKETS callables to run in order; after the i-th, the gradients of bucket i are final."""
        tape, marks = state["tape"], state["marks"]

        def prologue():
            self._grads = state["grads"]
            self._tape = None
            self._bwd_arena, self._bwd_used = torch.zeros(state["arena_numel"], device=self.dev, dtype=torch.float32), 0
            dur, B = state["dur"], state["B"]
            danc = torch.zeros_like(state["anc_s"])
            for s, (h, dp) in enumerate(zip(state["heads"], dpreds)):
                if dp is None:
                    continue
                dp = dp.contiguous().float()
                dh = self._grad(h)
                _lib.check(self.lib.yad_decode_bwd(h.ptr, h.ld, dp.data_ptr(), B, h.W, self.A, self.nc, state["anc_s"][s].data_ptr(),
                                                   float(state["strides"][s]) / float(state["center_scaler"]), dur, dh.ptr, dh.ld,
                                                   danc[s].data_ptr(), self._s()), "decode bwd")
            for s, a in enumerate(state["anchors"]):
                if a.requires_grad:
                    self._pgrad(a).add_(danc[s], alpha=dur)

        def run(lo, hi):
            for fn in reversed(tape[lo:hi]):
                fn()

        def seg0():
            prologue()
            run(marks[2], len(tape))          # neck, H-means, layer4

        def seg1():
            run(marks[1], marks[2])           # layer3

        def seg2():
            run(0, marks[1])                  # layer2, layer1, stem
            tape.clear()
            state["grads"].clear()
            self._grads, self._bwd_arena = {}, None
        return [seg0, seg1, seg2]

    def backward(self, state, dpreds: List[Optional[torch.Tensor]]):
        """Accumulates d loss / d parameter into every ``p.grad`` (allocated when missing)."""
        for i, seg in enumerate(self.backward_segments(state, dpreds)):
            seg()
            if self.on_bucket is not None:
                self.on_bucket(i)


class _TrainFn(torch.autograd.Function):
    """autograd hook: forward runs the kernels and keeps the tape; backward replays it and writes the parameter
    gradients straight into ``p.grad`` (the ``params`` inputs only exist so that autograd calls backward)."""

    @staticmethod
    def forward(ctx, eng: TrainEngine, xs: torch.Tensor, T: int, L_res: int, *params):
        record = any(ctx.needs_input_grad[4:])
        preds, state = eng.forward(xs, T, L_res, record)
        ctx.eng, ctx.state = eng, state
        return tuple(preds)

    @staticmethod
    def backward(ctx, *dpreds):
        if ctx.state is None or not ctx.state["tape"]:
            raise RuntimeError("yad_b200: backward through the same train-mode forward twice is not supported")
        with torch.no_grad():
            ctx.eng.backward(ctx.state, list(dpreds))
        return (None,) * len(ctx.needs_input_grad)


class _GraphFn(torch.autograd.Function):
    """Replays the captured forward / backward CUDA graphs of one (batch, length) shape."""

    @staticmethod
    def forward(ctx, g: dict, x: torch.Tensor, *params):
        g["x"].copy_(x)
        g["fwd"].replay()
        _lib.launch_count += g["fwd_launches"]
        ctx.g = g
        return tuple(p.detach() for p in g["preds"])

    @staticmethod
    def backward(ctx, *dpreds):
        g = ctx.g
        with torch.no_grad():
            for dst, dp in zip(g["dpreds"], dpreds):
                if dp is None:
                    dst.zero_()
                else:
                    dst.copy_(dp)
            hook = g["eng"].on_bucket
            for i, gb in enumerate(g["bwd"]):
                gb.replay()
                if hook is not None:
                    hook(i)
        _lib.launch_count += g["bwd_launches"]
        _lib.param_epoch += 1
        return (None,) * len(ctx.needs_input_grad)


def _capture_train_graphs(model, eng: TrainEngine, fe, x: torch.Tensor) -> dict:
    """Captures frontend + train-mode forward and the backward of one input shape into two CUDA graphs (one memory pool).
    The ~900 launches of a train step are CPU-bound when issued one by one; replayed they are GPU-bound."""
    B, _, L = x.shape
    L_res = -(-fe.rs_P * L // fe.rs_O)
    for p in eng._params:                      # the graphs write through these pointers: they must exist and stay put
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    g = {"x": torch.empty_like(x), "plan": {}, "fe": fe}      # `fe` owns the packed frontend constants the graph points into
    g["x"].copy_(x)
    eng._ensure_pack_table()                   # host -> device copy: must happen before the capture starts
    torch.cuda.synchronize(eng.dev)
    pool = torch.cuda.graph_pool_handle()
    fwd, bwd = torch.cuda.CUDAGraph(), [torch.cuda.CUDAGraph() for _ in range(TrainEngine.N_BUCKETS)]
    eng._force_repack = True                   # weight packing must be part of the graph (the parameters change every step)
    try:
        n0 = _lib.launch_count
        with torch.cuda.graph(fwd, pool=pool):
            xs = fe.run_frontend(g["x"], g["plan"], None, model._taper(g["x"]))
            preds, state = eng.forward(xs, xs.shape[-1], L_res, True)
        n1 = _lib.launch_count
        g["dpreds"] = [torch.zeros_like(p) for p in preds]
        torch.cuda.synchronize(eng.dev)
        # one graph per gradient bucket (same pool, replayed in capture order): the all-reduce of bucket i is launched between
        # replays and overlaps the later segments
        for gb, seg in zip(bwd, eng.backward_segments(state, g["dpreds"])):
            with torch.cuda.graph(gb, pool=pool):
                seg()
        n2 = _lib.launch_count
    finally:
        eng._force_repack = False
    _lib.launch_count = n0
    g["pack_table"] = eng._pack_table           # the captured pack kernel reads this device table: keep it alive with the graph
    g.update(eng=eng, fwd=fwd, bwd=bwd, preds=preds, state=state, fwd_launches=n1 - n0, bwd_launches=n2 - n1,
             grad_ptrs=[p.grad.data_ptr() for p in eng._params])
    return g


def run_train_forward(model, eng: TrainEngine, xs: torch.Tensor, T: int, L_res: int):
    eng._params = [p for p in model.parameters() if p.requires_grad]
    if torch.is_grad_enabled() and eng._params:
        return _TrainFn.apply(eng, xs, T, L_res, *eng._params)
    preds, _ = eng.forward(xs, T, L_res, False)
    return tuple(preds)


def run_train_forward_graphed(model, eng: TrainEngine, fe, x: torch.Tensor):
    """PCM -> predictions through CUDA graphs (``model.train_graphs = True``).  The first two calls of a shape run eagerly
    (lazy initialisation, allocator warm-up), the third captures, later ones replay.  Falls back to the eager path when a
    parameter's ``.grad`` buffer was replaced (``zero_grad(set_to_none=True)``) - graphs need stable pointers."""
    eng._params = [p for p in model.parameters() if p.requires_grad]
    key = (tuple(x.shape), tuple(p.data_ptr() for p in eng._params[:8]))
    ent = eng._graphs.get(key)
    if ent is None:
        ent = eng._graphs[key] = {"warm": 0, "g": None}
    if ent["g"] is None:
        if ent["warm"] < 2:
            ent["warm"] += 1
            return None
        ent["g"] = _capture_train_graphs(model, eng, fe, x)
    g = ent["g"]
    if any(p.grad is None or p.grad.data_ptr() != q for p, q in zip(eng._params, g["grad_ptrs"])):
        return None
    return _GraphFn.apply(g, x, *eng._params)
