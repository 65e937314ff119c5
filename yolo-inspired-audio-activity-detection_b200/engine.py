"""Inference engine: packs the model's weights for the sm_100a kernels and runs the whole forward
(frontend -> stem -> ResNet stages -> RepBi-PAN neck -> anchor decode) as a fixed list of C-ABI calls
on the current CUDA stream.

Data layout in HBM (per batch of B clips, T = frames = 960 for 60 s clips):
  pcm        [B, L]            f32   caller's tensor, read once
  mel        [B, 32, T]        f32   stage-A output (123 KB / clip)
  x_spectral [B, 2, 32, T]     f32   stage-B output (NCHW, as the reference's tensor)
  activations NHWC ("pixel-major"), bf16 (tcgen05 path) or f32 (CUDA-core parity path); channel
  pitches are multiples of 64 on the bf16 path (the 15-channel head tensors are zero-padded to 64) so
  every K-block of the implicit GEMM is one 128-byte swizzled row; torch.cat sites are channel slices
  of one buffer.
  preds      [B, 630, 3+nc]    f32

torch is used here for device memory (the workspace tensors) and the stream handle only.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
import weakref
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from . import frontend_consts as fc
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, BF16, F32, ConvDesc, FlatDesc


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def _fold_bn(w: torch.Tensor, b: Optional[torch.Tensor], bn: nn.BatchNorm2d) -> Tuple[torch.Tensor, torch.Tensor]:
    """conv -> eval-mode BatchNorm as one affine conv (fp32, done once at pack time)."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    wf = w * scale.reshape(-1, 1, 1, 1)
    b0 = b if b is not None else torch.zeros_like(bn.running_mean)
    return wf, (b0 - bn.running_mean) * scale + bn.bias


class _Conv:
    """One packed convolution (weights folded, both kernel layouts prepared lazily for the engine dtype)."""

    def __init__(self, name, w, b, stride, pad, act, dev, dtype, simt: bool = False):
        self.name = name
        self.simt = simt or dtype != BF16   # CUDA-core kernel ([kh][kw][Cin][Cout] weights): fp32 mode, or a bf16 conv whose
                                            # Cin is tiny (the 2-channel first conv of the custom backbone)
        self.cout, self.cin, self.kh, self.kw = w.shape
        self.sh, self.sw = (stride, stride) if isinstance(stride, int) else tuple(stride)
        self.ph, self.pw = (pad, pad) if isinstance(pad, int) else tuple(pad)
        self.act = act
        w = w.detach().to(dev, torch.float32)
        self.bias = b.detach().to(dev, torch.float32).contiguous()
        if dtype == BF16 and not simt:
            self.cin_pad = _ceil(self.cin, 64)
            self.cout_pad = _ceil(self.cout, 16)
            wt = torch.zeros(self.cout_pad, self.kh, self.kw, self.cin_pad, device=dev, dtype=torch.float32)
            wt[: self.cout, :, :, : self.cin] = w.permute(0, 2, 3, 1)
            self.w = wt.reshape(self.cout_pad, -1).to(torch.bfloat16).contiguous()
            # the same filter with its taps ordered (kw, kh): used when a conv is run with the roles of H and W swapped
            self.w_t = wt.permute(0, 2, 1, 3).reshape(self.cout_pad, -1).to(torch.bfloat16).contiguous() if self.kh * self.kw > 1 else self.w
            bp = torch.zeros(self.cout_pad, device=dev, dtype=torch.float32)
            bp[: self.cout] = self.bias
            self.bias = bp
        else:
            self.cin_pad = self.cin
            self.cout_pad = self.cout
            self.w = w.permute(2, 3, 1, 0).contiguous()     # [kh, kw, Cin, Cout]
            if dtype == BF16:
                self.w = self.w.to(torch.bfloat16)


class _SideHandle(C.c_void_p):
    """Stream argument of a call that runs on the engine's side stream: left alone when a recorded plan is replayed (the
    main-stream handles are patched to the caller's current stream)."""


class _RecLib:
    """The ctypes library seen through a thread-local recorder: while ``tls.rec`` is a list, every C-ABI call is appended to
    it as (function, argument list, name).  ``InferenceEngine.run`` records the first forward of a (batch, length) plan and
    replays the call list afterwards - the host side of a step then costs a few microseconds per launch instead of the Python
    that derives shapes, descriptors and buffers (which had become longer than the GPU step at 512 clips)."""

    def __init__(self, raw, tls):
        self.__dict__["_raw"] = raw
        self.__dict__["_tls"] = tls

    def __getattr__(self, name):
        fn = getattr(self._raw, name)
        tls = self._tls

        def call(*args):
            rc = fn(*args)
            rec = getattr(tls, "rec", None)
            if rec is not None:
                rec.append((fn, list(args), name))
            return rc
        self.__dict__[name] = call
        return call


class InferenceEngine:
    def __init__(self, model, device: torch.device, compute_dtype: str, frontend_only: bool = False):
        if device.type != "cuda":
            raise RuntimeError("yad_b200 needs the model on a CUDA (sm_100a) device; there is no CPU fallback")
        if compute_dtype not in ("bf16", "f32"):
            raise ValueError("compute_dtype must be 'bf16' (tcgen05 path) or 'f32' (CUDA-core parity path)")
        self.dev = device
        self._tls = threading.local()
        self.lib = _RecLib(_lib.init(device.index if device.index is not None else torch.cuda.current_device()), self._tls)
        self.dtype = BF16 if compute_dtype == "bf16" else F32
        self.use_graphs = os.environ.get("YAD_INFER_GRAPHS", "1") != "0"
        self.fused_neck = os.environ.get("YAD_FUSED_NECK", "1") != "0"
        self.dual_ds = os.environ.get("YAD_DUAL_DS", "1") != "0"
        self.s2d_route = os.environ.get("YAD_S2D", "1") != "0"
        self.fold_ds = os.environ.get("YAD_FOLD_DS", "1") != "0"
        self.stem_overlap = os.environ.get("YAD_STEM_OVERLAP", "1") != "0"
        self._threads_seen = set()
        self.tdtype = torch.bfloat16 if self.dtype == BF16 else torch.float32
        self.cfg = model.config
        self.nc = model.num_classes
        self.A = self.cfg["num_anchors"]
        self.E = 3 + self.nc
        self.n_head = self.A * self.E
        with torch.no_grad():
            self._pack_frontend(model)
            if frontend_only:      # train mode: the CNN runs in train_engine.py on the live parameters
                return
            self._pack_cnn(model)
            dur = float(self.cfg["sample_duration"])
            self.anchors_s = torch.cat([model.sm_anchors * dur, model.md_anchors * dur, model.lg_anchors * dur]).float().cpu()

    # ------------------------------------------------------------------ packing
    def _pack_frontend(self, model):
        dev = self.dev
        k = model.resampler.kernel
        g = math.gcd(int(self.cfg["sample_rate"]), int(self.cfg["new_sample_rate"]))
        self.rs_O = int(self.cfg["sample_rate"]) // g
        P = int(self.cfg["new_sample_rate"]) // g
        pk = fc.pack_resample_taps(k, orig_step=self.rs_O, hops_per_group=(8 * 1000) // P if (8 * 1000) % P == 0 else 0)
        self.rs_P, self.rs_KW = pk["P"], pk["KW"]
        if int(self.cfg["new_sample_rate"]) // g != self.rs_P:
            raise ValueError("resampler.kernel does not match sample_rate/new_sample_rate")
        self.rs_width = (self.rs_KW - self.rs_O) // 2
        self.rs_taps = pk["taps"].to(dev)
        self.rs_base = pk["base"].to(dev)
        self.rs_lane_map = None if pk["lane_map"] is None else pk["lane_map"].to(dev)
        self.rs_window_len = pk["window_len"]
        self.win = model.melspectogram_tfmr.spectrogram.window.detach().to(dev, torch.float32).contiguous()
        csr = fc.pack_mel_csr(model.melspectogram_tfmr.mel_scale.fb)
        self.fb_val, self.fb_bin, self.fb_start = csr["val"].to(dev), csr["bin"].to(dev), csr["start"].to(dev)
        self.dct = model.mfcc_tfmr.dct_mat.detach().to(dev, torch.float32).contiguous()
        self.tw = fc.fft_twiddles(1000).to(dev)

    def _cbl(self, name, m):  # ConvBorINorm
        w, b = _fold_bn(m.conv.weight, m.conv.bias, m.norm)
        return _Conv(name, w, b, m.conv.stride, m.conv.padding, ACT_LRELU if m.has_activation else ACT_NONE, self.dev, self.dtype)

    def _rep(self, name, m):
        """RepVGG block -> either one folded conv (deploy) or the three train-form branches."""
        if hasattr(m, "conv_reparam"):
            c = m.conv_reparam
            return {"deploy": _Conv(name, c.weight, c.bias, 1, 1, ACT_LRELU, self.dev, self.dtype)}
        out = {"c3": self._cbl(name + ".conv3x3", m.conv3x3), "c1": self._cbl(name + ".conv1x1", m.conv1x1)}
        if isinstance(m.identity, nn.BatchNorm2d):
            bn = m.identity
            s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            out["id_scale"] = s.detach().to(self.dev, torch.float32).contiguous()
            out["id_shift"] = (bn.bias - bn.running_mean * s).detach().to(self.dev, torch.float32).contiguous()
        return out

    def _pack_custom(self, fe):
        """CustomBackBone (modules/_backbone.py:82-116): first_conv + 5 ExtractorBlocks of ExtractorLayers (:8-46)."""
        dev = self.dev
        w, b = _fold_bn(fe.first_conv[0].weight, fe.first_conv[0].bias, fe.first_conv[1])
        self.first_conv = _Conv("fe.first_conv", w, b, 1, 3, ACT_LRELU, dev, self.dtype, simt=True)
        self.cblocks = []
        for bn_ in ("entry_block", "block1", "block2", "block3", "block4"):
            layers = []
            for ln, lay in getattr(fe, bn_).module_dict.items():
                ca, bna, cb, bnb = lay._layer[0], lay._layer[1], lay._layer[3], lay._layer[4]
                wa, ba = _fold_bn(ca.weight, ca.bias, bna)
                wb, bb = _fold_bn(cb.weight, cb.bias, bnb)
                r = lay._res_layer
                layers.append({
                    "a": _Conv(f"{bn_}.{ln}.a", wa, ba, ca.stride, ca.padding, ACT_LRELU, dev, self.dtype),
                    "b": _Conv(f"{bn_}.{ln}.b", wb, bb, cb.stride, cb.padding, ACT_NONE, dev, self.dtype),
                    "r": _Conv(f"{bn_}.{ln}.res", r.weight, r.bias, r.stride, r.padding, ACT_NONE, dev, self.dtype),
                })
            self.cblocks.append(layers)

    def _pack_cnn(self, model):
        fe, ms = model.feature_extractor, model.multiscale_module
        dev = self.dev
        self.bb_kind = "custom" if hasattr(fe, "first_conv") else ("bottleneck" if fe.block_name == "Bottleneck" else "basic")
        if self.bb_kind == "custom":
            self._pack_custom(fe)
        else:
            self._pack_resnet(fe)
        self._pack_neck(ms)

    def _pack_resnet(self, fe):
        dev = self.dev
        # stem conv1: raw (no BN / bias / activation before conv2, modules/_backbone.py:143-144); [7][7][2][64] f32
        self.stem_w = fe.conv1.weight.detach().to(dev, torch.float32).permute(2, 3, 1, 0).contiguous()
        if self.dtype == BF16:
            # tensor-core stem: [64][K=112] bf16, K = kh*16 + kw*2 + c, stored as 8x8 core matrices (no-swizzle K-major)
            wk = torch.zeros(64, 7, 16, device=dev, dtype=torch.float32)
            wk[:, :, :14] = fe.conv1.weight.detach().to(dev, torch.float32).permute(0, 2, 3, 1).reshape(64, 7, 14)
            self.stem_w_tc = wk.reshape(8, 8, 14, 8).permute(0, 2, 1, 3).contiguous().to(torch.bfloat16)
        w2, b2 = _fold_bn(fe.conv2.weight, None, fe.bn1)
        self.conv2 = _Conv("fe.conv2", w2, b2, 2, 3, ACT_RELU, dev, self.dtype)
        if self.dtype == BF16 and self.bb_kind == "basic":
            # conv2 over the space-to-depth stem output: tap (kh, kw) reads parity plane ((kh-3)&1, (kw-3)&1) shifted by
            # (floor((kh-3)/2), floor((kw-3)/2)); steps grouped by plane
            steps = []
            for kh in range(7):
                for kw in range(7):
                    a, b_ = (kh - 3) // 2, (kw - 3) // 2
                    steps.append((((kh - 3) - 2 * a) * 2 + ((kw - 3) - 2 * b_), a, b_, (kh * 7 + kw) * 64))
            steps.sort(key=lambda t: t[0])
            arr = lambda i: (C.c_int32 * len(steps))(*[t[i] for t in steps])   # noqa: E731
            self.conv2_steps = (arr(0), arr(1), arr(2), arr(3))
        self.fused_stem = (self.dtype == BF16 and self.bb_kind == "basic" and os.environ.get("YAD_FUSED_STEM", "1") != "0")
        if self.fused_stem:
            self._pack_fused_stem(fe.conv1.weight, w2, b2)
        self.stages = []
        for li in range(1, 5):
            blocks = []
            for bi, blk in enumerate(getattr(fe, f"layer{li}")):
                if hasattr(blk, "conv3"):      # torchvision Bottleneck ([tv] models/resnet.py:143-163): the 3x3 carries the stride
                    w1, b1 = _fold_bn(blk.conv1.weight, None, blk.bn1)
                    w2, b2_ = _fold_bn(blk.conv2.weight, None, blk.bn2)
                    w3, b3 = _fold_bn(blk.conv3.weight, None, blk.bn3)
                    ent = {
                        "c1": _Conv(f"l{li}.{bi}.conv1", w1, b1, 1, 0, ACT_RELU, dev, self.dtype),
                        "c2": _Conv(f"l{li}.{bi}.conv2", w2, b2_, blk.stride, 1, ACT_RELU, dev, self.dtype),
                        "c3": _Conv(f"l{li}.{bi}.conv3", w3, b3, 1, 0, ACT_RELU, dev, self.dtype),   # relu after the residual add
                    }
                    if blk.downsample is not None:
                        wd, bd = _fold_bn(blk.downsample[0].weight, None, blk.downsample[1])
                        ent["ds"] = _Conv(f"l{li}.{bi}.ds", wd, bd, blk.stride, 0, ACT_NONE, dev, self.dtype)
                    blocks.append(ent)
                    continue
                w1, b1 = _fold_bn(blk.conv1.weight, None, blk.bn1)
                wB, bB = _fold_bn(blk.conv2.weight, None, blk.bn2)
                ent = {
                    "c1": _Conv(f"l{li}.{bi}.conv1", w1, b1, blk.stride, 1, ACT_RELU, dev, self.dtype),
                    "c2": _Conv(f"l{li}.{bi}.conv2", wB, bB, 1, 1, ACT_RELU, dev, self.dtype),  # relu after the residual add
                }
                if blk.downsample is not None:
                    wd, bd = _fold_bn(blk.downsample[0].weight, None, blk.downsample[1])
                    ent["ds"] = _Conv(f"l{li}.{bi}.ds", wd, bd, blk.stride, 0, ACT_NONE, dev, self.dtype)
                blocks.append(ent)
            self.stages.append(blocks)

    # column tap sets of conv2 that stay inside conv1's output: (lo, hi) = valid kw2 range; index = variant of the fix-up weights
    _COL_VARIANTS = [(3, 6), (1, 6), (0, 3), (0, 4), (0, 5)]
    _ROW_CLASSES = [(0, 6), (3, 6), (1, 6), (0, 4)]        # rows 2..6 | row 0 | row 1 | row 7 (H = 32 -> 16 -> 8)

    def _pack_fused_stem(self, w1, w2f, b2):
        """conv1 o (conv2 * bn1) as one 19x19 stride-4 convolution (csrc/conv_stem_fused.cu): composite weights
        W12[co, c, dh, dw] = sum_{c1, kh2, kw2 valid} W2[co, c1, kh2, kw2] W1[c1, c, dh - 2 kh2, dw - 2 kw2], built in fp64 as a
        transposed convolution; one variant per set of conv2 taps that fall inside conv1's output (zero padding of the
        INTERMEDIATE tensor, modules/_backbone.py:143-144)."""
        dev = self.dev
        w1d = w1.detach().to(dev, torch.float64)
        w2d = w2f.detach().to(dev, torch.float64)

        def compose(rows, cols):
            m = torch.zeros_like(w2d)
            m[:, :, rows[0]:rows[1] + 1, cols[0]:cols[1] + 1] = w2d[:, :, rows[0]:rows[1] + 1, cols[0]:cols[1] + 1]
            return torch.nn.functional.conv_transpose2d(m, w1d, stride=2)            # [64, 2, 19, 19]
        cls = []
        for rows in self._ROW_CLASSES:
            w12 = compose(rows, (0, 6)).permute(0, 2, 3, 1).reshape(64, 19, 38)      # K order (dh, dw, c)
            wk = torch.zeros(64, 19, 48, device=dev, dtype=torch.float64)
            wk[:, :, :38] = w12
            cls.append(wk.reshape(8, 8, 114, 8).permute(0, 2, 1, 3).contiguous())    # [n/8][k/8][n%8][k%8]
        self.fstem_w = torch.stack(cls).to(torch.bfloat16).contiguous()
        var = [torch.stack([compose(rows, cols).permute(2, 3, 1, 0) for rows in self._ROW_CLASSES]) for cols in self._COL_VARIANTS]
        self.fstem_wvar = torch.stack(var).to(torch.float32).contiguous()            # [5][4][19][19][2][64]
        self.fstem_bias = b2.detach().to(dev, torch.float32).contiguous()

    @classmethod
    def _fused_stem_border_cols(cls, W: int):
        """Output columns whose conv2 taps partly fall outside conv1's output, with their fix-up weight variant."""
        W1 = (W - 1) // 2 + 1
        Wo = (W1 - 1) // 2 + 1
        cols, var = [], []
        for wo in list(range(0, min(2, Wo))) + list(range(max(Wo - 3, 2), Wo)):
            v = [k for k in range(7) if 0 <= 2 * wo - 3 + k < W1]
            if len(v) == 7:
                continue
            cols.append(wo)
            var.append(cls._COL_VARIANTS.index((v[0], v[-1])))
        return cols, var

    def _pack_neck(self, ms):
        sp = ms.cspsppf
        self.n = {
            "sp1": self._cbl("sp.c1", sp.conv_1_3_4[0]), "sp3": self._cbl("sp.c3", sp.conv_1_3_4[1]),
            "sp4": self._cbl("sp.c4", sp.conv_1_3_4[2]), "sp2": self._cbl("sp.c2", sp.conv2),
            "sp5": self._cbl("sp.c5", sp.conv5), "sp6": self._cbl("sp.c6", sp.conv6), "sp7": self._cbl("sp.c7", sp.conv7),
            "b3c1": self._cbl("bic3.c1", ms.bic3.conv_c1), "b3c0": self._cbl("bic3.c0", ms.bic3.conv_c0),
            "b3o": self._cbl("bic3.out", ms.bic3.conv_out),
            "b2c1": self._cbl("bic2.c1", ms.bic2.conv_c1), "b2c0": self._cbl("bic2.c0", ms.bic2.conv_c0),
            "b2o": self._cbl("bic2.out", ms.bic2.conv_out),
            "ds2": self._cbl("conv2_ds", ms.conv2_downsample), "ds3": self._cbl("conv3_ds", ms.conv3_downsample),
        }
        self.rep = {}
        for nm in ("rep_block2_1", "rep_block3_1", "rep_block3_2", "rep_block4_1"):
            rb = getattr(ms, nm)
            blocks = [rb.conv1] + (list(rb.blocks) if isinstance(rb.blocks, nn.Sequential) else [])
            self.rep[nm] = [self._rep(f"{nm}.{i}", b) for i, b in enumerate(blocks)]

    def _fused_neck_for(self, geoms, fmaps):
        """The compiled fused-neck program of this feature-map geometry (None when the geometry / model form is not covered:
        the layer-by-layer path runs instead)."""
        key = tuple(geoms)
        cache = self.__dict__.setdefault("_fused_necks", {})
        if key not in cache:
            from .neck_fused import FusedNeck
            try:
                cache[key] = FusedNeck(self, [g_[0] for g_ in geoms], [g_[1] for g_ in geoms], [f.shape[3] for f in fmaps])
            except (NotImplementedError, ValueError):
                cache[key] = None
        return cache[key]

    # ------------------------------------------------------------------ shapes
    def frames(self, L: int) -> int:
        target = -(-self.rs_P * L // self.rs_O)
        return target // 1000

    def grids(self, L: int) -> List[int]:
        """Grid sizes of the three heads for an input of L samples (T/8, T/16, T/32 for the default net)."""
        T = self.frames(L)
        w = T
        ws = []
        for _ in range(2):
            w = (w + 6 - 7) // 2 + 1
        ws.append(w)                       # layer1
        for _ in range(3):
            w = (w + 2 - 3) // 2 + 1
            ws.append(w)
        return [ws[1], ws[2], ws[3]]

    # ------------------------------------------------------------------ execution helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    # ---- a second stream for work that is independent of the main chain (the stem's border-column fix-up): fork / join are
    #      recorded into the plan as pseudo-calls, so a replay - and a CUDA graph captured from it - keeps the two branches
    def _side_stream(self) -> torch.cuda.Stream:
        st = getattr(self._tls, "side", None)
        if st is None:
            st = self._tls.side = torch.cuda.Stream(self.dev)
            self._tls.ev_fork, self._tls.ev_join = torch.cuda.Event(), torch.cuda.Event()
        return st

    def _fork(self) -> int:
        side = self._side_stream()
        self._tls.ev_fork.record(torch.cuda.current_stream(self.dev))
        side.wait_event(self._tls.ev_fork)
        return 0

    def _join(self) -> int:
        self._tls.ev_join.record(self._side_stream())
        torch.cuda.current_stream(self.dev).wait_event(self._tls.ev_join)
        return 0

    def _pseudo(self, fn, name: str) -> None:
        fn()
        rec = getattr(self._tls, "rec", None)
        if rec is not None:
            rec.append((fn, [], name))

    def _buf(self, plan, name, B, H, W, ld, zero=False, dtype=None):
        t = plan.get(name)
        if t is None:
            f = torch.zeros if zero else torch.empty
            t = f((B, H, W, ld), device=self.dev, dtype=dtype or self.tdtype)
            plan[name] = t
        return t

    def _conv(self, cv: _Conv, x: torch.Tensor, cin_off: int, out: torch.Tensor, co_off: int, res: Optional[torch.Tensor] = None,
              act: Optional[int] = None, out2: Optional[torch.Tensor] = None):
        """x [B,H,W,ld_in] (channels [cin_off, cin_off+Cin) are read), out [B,Ho,Wo,ld_out] slice at co_off."""
        B, H, W, ld_in = x.shape
        es = x.element_size()
        d = ConvDesc(B=B, H=H, W=W, Cin=cv.cin_pad if self.dtype == BF16 else cv.cin, ld_in=ld_in, Cout=cv.cout,
                     ld_out=out.shape[3], co_off=co_off, kh=cv.kh, kw=cv.kw, sh=cv.sh, sw=cv.sw, ph=cv.ph, pw=cv.pw,
                     act=cv.act if act is None else act, ld_res=0 if res is None else res.shape[3])
        Ho = (H + 2 * cv.ph - cv.kh) // cv.sh + 1
        Wo = (W + 2 * cv.pw - cv.kw) // cv.sw + 1
        assert out.shape[0] == B and out.shape[1] == Ho and out.shape[2] == Wo, (cv.name, tuple(out.shape), (B, Ho, Wo))
        in_ptr = x.data_ptr() + cin_off * es
        if not x.is_contiguous():      # strided pixel view (e.g. the H = 1 stage read straight out of its flat layout)
            assert x.stride(3) == 1 and all(st % ld_in == 0 for st in x.stride()[:3]), (cv.name, x.stride())
            d.in_sw, d.in_sh, d.in_sb = x.stride(2) // ld_in, x.stride(1) // ld_in, x.stride(0) // ld_in
            assert self.dtype == BF16 and not cv.simt, "strided inputs are a tensor-core path feature"
        if self.dtype == BF16 and not cv.simt:
            assert cin_off + cv.cin_pad <= ld_in, (cv.name, cin_off, cv.cin_pad, ld_in)
            rc = self.lib.yad_conv_tc(C.byref(d), in_ptr, cv.w.data_ptr(), cv.cout_pad, cv.bias.data_ptr(), _lib.ptr(res),
                                      out.data_ptr(), BF16, _lib.ptr(out2), 0 if out2 is None else out2.shape[3], self._stream())
        else:
            d.Cin = cv.cin
            rc = self.lib.yad_conv_simt(C.byref(d), self.dtype, in_ptr, cv.w.data_ptr(), cv.cout, cv.bias.data_ptr(), _lib.ptr(res),
                                        out.data_ptr(), self._stream())
            if out2 is not None:
                raise AssertionError("out2 is only used on the tensor-core path")
        _lib.check(rc, f"conv {cv.name}")

    # ---- flat (halo-padded, h-fastest) layout of the bf16 backbone: tensor [B, Wp, Hp, ld], see include/yad_b200.h
    @staticmethod
    def _flat_geom(H: int, W: int) -> Tuple[int, int]:
        return (H + 1 if H > 1 else 1), W + 1

    def _flat_buf(self, plan, name, B, H, W, ld):
        t = plan.get(name)
        if t is None:
            Hp, Wp = self._flat_geom(H, W)
            t = plan[name] = torch.zeros((B, Wp, Hp, ld), device=self.dev, dtype=torch.bfloat16)   # halo stays zero forever
        return t

    def _conv_flat(self, cv: _Conv, x: torch.Tensor, H: int, W: int, out: torch.Tensor, res: Optional[torch.Tensor] = None):
        """Stride-1 'same' conv, flat in -> flat out (same geometry), patch-resident tcgen05 kernel."""
        B, Wp, Hp, ld_in = x.shape
        assert out.shape[:3] == x.shape[:3] and cv.sh == 1 and cv.sw == 1
        d = FlatDesc(B=B, H=H, W=W, Hp=Hp, Wp=Wp, Cin=cv.cin_pad, ld_in=ld_in, Cout=cv.cout, ld_out=out.shape[3], co_off=0,
                     kh=cv.kh, kw=cv.kw, ph=cv.ph, pw=cv.pw, act=cv.act, ld_res=0 if res is None else res.shape[3])
        rc = self.lib.yad_conv_flat(C.byref(d), x.data_ptr(), cv.w.data_ptr(), cv.cout_pad, cv.bias.data_ptr(), _lib.ptr(res),
                                    out.data_ptr(), 0, self._stream())
        _lib.check(rc, f"conv_flat {cv.name}")

    def _conv_flat_with_s2d(self, cv: _Conv, x: torch.Tensor, H: int, W: int, out: torch.Tensor, res: Optional[torch.Tensor],
                            s2d: torch.Tensor):
        """_conv_flat that also writes the space-to-depth copy the next layer's stride-2 block reads (yad_conv_flat_s2d)."""
        B, Wp, Hp, ld_in = x.shape
        _, Wp2, Hp2, ld2 = s2d.shape
        assert out.shape[:3] == x.shape[:3] and cv.sh == 1 and cv.sw == 1 and ld2 == 4 * cv.cout
        d = FlatDesc(B=B, H=H, W=W, Hp=Hp, Wp=Wp, Cin=cv.cin_pad, ld_in=ld_in, Cout=cv.cout, ld_out=out.shape[3], co_off=0,
                     kh=cv.kh, kw=cv.kw, ph=cv.ph, pw=cv.pw, act=cv.act, ld_res=0 if res is None else res.shape[3])
        rc = self.lib.yad_conv_flat_s2d(C.byref(d), x.data_ptr(), cv.w.data_ptr(), cv.cout_pad, cv.bias.data_ptr(), _lib.ptr(res),
                                        out.data_ptr(), s2d.data_ptr(), Hp2, Wp2, self._stream())
        _lib.check(rc, f"conv_flat (+ space-to-depth copy) {cv.name}")

    def _s2d_steps(self, cv: _Conv, H: int, W: int, Ho: int, Wo: int):
        """(chunk, dh, dw, wk) step lists of a stride-2 convolution read from the space-to-depth copy of its input: tap (kh, kw)
        reads parity plane ((kh - ph) & 1, (kw - pw) & 1) shifted by (floor((kh - ph) / 2), floor((kw - pw) / 2)); taps that only
        ever see padding are dropped; steps grouped by ascending 64-channel chunk (one patch load per plane chunk)."""
        key = (cv.name, H, W)
        cache = self.__dict__.setdefault("_s2d_step_cache", {})
        if key not in cache:
            nch = cv.cin_pad // 64
            steps = []
            for kh in range(cv.kh):
                rh = kh - cv.ph
                if not any(0 <= 2 * h2 + rh < H for h2 in range(Ho)):
                    continue
                for kw in range(cv.kw):
                    rw = kw - cv.pw
                    if not any(0 <= 2 * w2 + rw < W for w2 in range(Wo)):
                        continue
                    a, b_ = rh // 2, rw // 2
                    plane = (rh - 2 * a) * 2 + (rw - 2 * b_)
                    for c in range(nch):
                        steps.append((plane * nch + c, a, b_, (kh * cv.kw + kw) * cv.cin_pad + c * 64))
            steps.sort(key=lambda t: t[0])
            arr = lambda i: (C.c_int32 * len(steps))(*[t[i] for t in steps])   # noqa: E731
            cache[key] = (arr(0), arr(1), arr(2), arr(3), len(steps))
        return cache[key]

    def _conv_s2d(self, cv: _Conv, s2d: torch.Tensor, H: int, W: int, out: torch.Tensor, Ho: int, Wo: int):
        """Stride-2 convolution (3x3 pad 1, or the 1x1 pad 0 downsample) of the layer whose output was also written in
        space-to-depth form: a stride-1 flat convolution over (plane, shift) steps in the patch-resident kernel."""
        B, Wp2, Hp2, ld_in = s2d.shape
        assert out.shape[1:3] == (Wp2, Hp2) and ld_in == 4 * cv.cin_pad and (cv.sh, cv.sw) == (2, 2)
        ch, dh, dw, wk, n = self._s2d_steps(cv, H, W, Ho, Wo)
        d = FlatDesc(B=B, H=Ho, W=Wo, Hp=Hp2, Wp=Wp2, Cin=ld_in, ld_in=ld_in, Cout=cv.cout, ld_out=out.shape[3], co_off=0, kh=1, kw=1,
                     ph=0, pw=0, act=cv.act, ld_res=0, Hp_out=Hp2, Wp_out=Wp2)
        rc = self.lib.yad_conv_flat_taps(C.byref(d), n, ch, dh, dw, wk, cv.kh * cv.kw * cv.cin_pad, s2d.data_ptr(), cv.w.data_ptr(),
                                         cv.cout_pad, cv.bias.data_ptr(), 0, out.data_ptr(), self._stream())
        _lib.check(rc, f"conv (stride 2, space-to-depth) {cv.name}")

    def _conv_flat_plus_ds(self, blk: dict, t: torch.Tensor, s2d: torch.Tensor, H: int, W: int, out: torch.Tensor):
        """out = relu(conv2(t) + downsample(x)) as ONE flat convolution (yad_conv_flat_taps2): the 1x1 stride-2 downsample of x is
        read from plane 0 of x's space-to-depth copy as Cin(x) / 64 extra K steps; weights concatenated along K and biases summed
        once per engine (torchvision BasicBlock, resnet.py:96-103)."""
        cv, ds = blk["c2"], blk["ds"]
        B, Wp, Hp, ld_in = t.shape
        assert s2d.shape[:3] == t.shape[:3] and out.shape[:3] == t.shape[:3] and ds.cout_pad == cv.cout_pad and cv.kh == 3 and cv.ph == 1
        fz = blk.get("c2ds")
        if fz is None:
            fz = blk["c2ds"] = {"w": torch.cat([cv.w, ds.w], 1).contiguous(), "bias": (cv.bias + ds.bias).contiguous(), "steps": {}}
        st = fz["steps"].get((H, W))
        if st is None:
            steps = []
            for c in range(cv.cin_pad // 64):
                for kh in range(3):
                    for kw in range(3):
                        if abs(kh - 1) < H and abs(kw - 1) < W:
                            steps.append((0, c, kh - 1, kw - 1, (kh * 3 + kw) * cv.cin_pad + c * 64))
            for c in range(ds.cin_pad // 64):
                steps.append((1, c, 0, 0, 9 * cv.cin_pad + c * 64))
            arr = lambda i: (C.c_int32 * len(steps))(*[s_[i] for s_ in steps])   # noqa: E731
            st = fz["steps"][(H, W)] = (arr(0), arr(1), arr(2), arr(3), arr(4), len(steps))
        d = FlatDesc(B=B, H=H, W=W, Hp=Hp, Wp=Wp, Cin=cv.cin_pad, ld_in=ld_in, Cout=cv.cout, ld_out=out.shape[3], co_off=0, kh=1, kw=1,
                     ph=0, pw=0, act=cv.act, ld_res=0)
        rc = self.lib.yad_conv_flat_taps2(C.byref(d), st[5], st[0], st[1], st[2], st[3], st[4], 9 * cv.cin_pad + ds.cin_pad, t.data_ptr(),
                                          s2d.data_ptr(), ds.cin_pad, s2d.shape[3], fz["w"].data_ptr(), cv.cout_pad, fz["bias"].data_ptr(), 0,
                                          out.data_ptr(), self._stream())
        _lib.check(rc, f"conv_flat (+ downsample as K steps) {cv.name}")

    def _conv_flat_s2(self, cv: _Conv, x: torch.Tensor, H: int, W: int, out: torch.Tensor, Ho: int, Wo: int,
                      ds: Optional[_Conv] = None, out_ds: Optional[torch.Tensor] = None):
        """Strided conv, flat in -> flat out: the tap-by-tap kernel run with the roles of H and W swapped (so that the
        input pixel strides are ascending for its 4-D tensor map), transposed filter taps.  With ``ds`` the BasicBlock's 1x1
        downsample branch of the same input is computed in the same launch (yad_conv_tc_dual) into ``out_ds``."""
        B, Wp, Hp, ld_in = x.shape
        _, Wpo, Hpo, ld_out = out.shape
        d = ConvDesc(B=B, H=W, W=H, Cin=cv.cin_pad, ld_in=ld_in, Cout=cv.cout, ld_out=ld_out, co_off=0, kh=cv.kw, kw=cv.kh,
                     sh=cv.sw, sw=cv.sh, ph=cv.pw, pw=cv.ph, act=cv.act, ld_res=0,
                     in_sw=1, in_sh=Hp, in_sb=Wp * Hp, out_sw=1, out_sh=Hpo, out_sb=Wpo * Hpo)
        assert (W + 2 * cv.pw - cv.kw) // cv.sw + 1 == Wo and (H + 2 * cv.ph - cv.kh) // cv.sh + 1 == Ho
        if ds is not None:
            assert (ds.kh, ds.kw, ds.ph, ds.pw) == (1, 1, 0, 0) and (ds.sh, ds.sw) == (cv.sh, cv.sw) and ds.cin_pad == cv.cin_pad
            assert ds.cout_pad == cv.cout_pad == cv.cout and out_ds.shape == out.shape
            rc = self.lib.yad_conv_tc_dual(C.byref(d), x.data_ptr(), cv.w_t.data_ptr(), cv.cout_pad, cv.bias.data_ptr(), out.data_ptr(),
                                           ds.w_t.data_ptr(), ds.bias.data_ptr(), ds.act, out_ds.data_ptr(), self._stream())
            _lib.check(rc, f"conv(s2, flat, + downsample) {cv.name}")
            return
        rc = self.lib.yad_conv_tc(C.byref(d), x.data_ptr(), cv.w_t.data_ptr(), cv.cout_pad, cv.bias.data_ptr(), 0, out.data_ptr(),
                                  BF16, 0, 0, self._stream())
        _lib.check(rc, f"conv(s2, flat) {cv.name}")

    def _repblock(self, plan, key, blocks, x, cin_off, out, co_off, out2=None):
        """RepBlock: chain of RepVGG blocks; the last one writes (out, co_off) [and the fp32 copy out2]."""
        B, H, W, _ = x.shape
        cur, cur_off = x, cin_off
        for i, blk in enumerate(blocks):
            last = i == len(blocks) - 1
            cout = (blk["deploy"] if "deploy" in blk else blk["c3"]).cout
            ld = _ceil(cout, 64) if self.dtype == BF16 else cout
            dst, dst_off = (out, co_off) if last else (self._buf(plan, f"{key}.t{i}", B, H, W, ld, zero=True), 0)
            if "deploy" in blk:
                self._conv(blk["deploy"], cur, cur_off, dst, dst_off, out2=out2 if last else None)
            else:
                a = self._buf(plan, f"{key}.a{i}", B, H, W, ld, zero=True)
                b = self._buf(plan, f"{key}.b{i}", B, H, W, ld, zero=True)
                self._conv(blk["c3"], cur, cur_off, a, 0)
                self._conv(blk["c1"], cur, cur_off, b, 0)
                has_id = "id_scale" in blk
                es = cur.element_size()
                rc = self.lib.yad_repvgg_merge(a.data_ptr(), b.data_ptr(), (cur.data_ptr() + cur_off * es) if has_id else 0,
                                               _lib.ptr(blk.get("id_scale")), _lib.ptr(blk.get("id_shift")), self.dtype,
                                               B * H * W, cout, ld, cur.shape[3], dst.data_ptr(), dst.shape[3], dst_off,
                                               ACT_LRELU, self._stream())
                _lib.check(rc, f"repvgg_merge {key}.{i}")
                if last and out2 is not None:
                    out2.copy_(dst[..., dst_off:dst_off + out2.shape[3]])
                    self._tls.rec = None      # a torch op sits between the C-ABI calls: this plan is not replayable
            cur, cur_off = dst, dst_off

    # ------------------------------------------------------------------ forward
    def run_frontend(self, x: torch.Tensor, plan: dict, taps: Optional[dict] = None, taper: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, _, L = x.shape
        T = self.frames(L)
        if T < 32:
            raise ValueError(f"input too short: {L} samples give {T} frames (need >= 32)")
        i16 = x.dtype == torch.int16          # 16-bit PCM (file sample format): converted (x / 32768) inside the kernel
        x = x.contiguous() if i16 else x.contiguous().float()
        mel = plan.get("mel")
        if mel is None:
            mel = plan["mel"] = torch.empty((B, 32, T), device=self.dev, dtype=torch.float32)
            plan["xs"] = torch.empty((B, 2, 32, T), device=self.dev, dtype=torch.float32)
        xs = plan["xs"]
        if taper is not None:       # taper_input: true (modules/_architecture.py:87-94)
            if taper.device != self.dev or taper.dtype != torch.float32 or not taper.is_contiguous():
                raise RuntimeError("taper_window must be a contiguous fp32 tensor on the model's device")
            fn = lambda *a: self.lib.yad_frontend_mel_power_taper(a[0], 1 if i16 else 0, taper.data_ptr(), taper.numel(), *a[1:])   # noqa: E731
        else:
            fn = self.lib.yad_frontend_mel_power_i16 if i16 else self.lib.yad_frontend_mel_power
        rc = fn(x.data_ptr(), B, L, self.rs_P, self.rs_O, self.rs_width, self.rs_taps.data_ptr(),
                                             self.rs_base.data_ptr(), _lib.ptr(self.rs_lane_map), self.rs_window_len, self.win.data_ptr(), self.tw.data_ptr(),
                                             self.fb_val.data_ptr(), self.fb_bin.data_ptr(), self.fb_start.data_ptr(),
                                             self.fb_val.numel(), mel.data_ptr(), T, self._stream())
        _lib.check(rc, "frontend_mel_power")
        tm = {}
        if taps is not None:
            for k in ("meldb", "mfcc", "mfdb"):
                tm[k] = torch.empty((B, 1, 32, T), device=self.dev, dtype=torch.float32)
        if getattr(self, "fused_stem", False) and T <= 1024:
            # x_spectral a second time as padded channel-interleaved bf16 words: the patch rows of the fused stem (bulk copies)
            xb = plan.get("xs_bf16")
            if xb is None:
                W1 = (T - 1) // 2 + 1
                n_seg = (((W1 - 1) // 2 + 1) + 127) // 128
                pitch = _ceil(max(512 * (n_seg - 1) + 536, 9 + T), 4)
                xb = plan["xs_bf16"] = torch.zeros((B, 32, pitch), device=self.dev, dtype=torch.int32)
            rc = self.lib.yad_frontend_finish_bf16(mel.data_ptr(), B, T, self.dct.data_ptr(), 80.0, 1 if self.cfg["scale_input"] else 0,
                                                   xs.data_ptr(), _lib.ptr(tm.get("meldb")), _lib.ptr(tm.get("mfcc")),
                                                   _lib.ptr(tm.get("mfdb")), xb.data_ptr(), xb.shape[2], 9, self._stream())
        else:
            rc = self.lib.yad_frontend_finish(mel.data_ptr(), B, T, self.dct.data_ptr(), 80.0, 1 if self.cfg["scale_input"] else 0,
                                              xs.data_ptr(), _lib.ptr(tm.get("meldb")), _lib.ptr(tm.get("mfcc")),
                                              _lib.ptr(tm.get("mfdb")), self._stream())
        _lib.check(rc, "frontend_finish")
        if taps is not None:
            taps.update(tm)
            taps["mel"] = mel.reshape(B, 1, 32, T).clone()
            taps["x_spectral"] = xs.clone()
        return xs

    def run_cnn(self, xs: torch.Tensor, plan: dict, taps: Optional[dict] = None) -> List[torch.Tensor]:
        """x_spectral [B,2,32,T] f32 -> three head tensors [B,G,ld] f32 (first n_head channels valid)."""
        B, _, H0, T = xs.shape
        bf = self.dtype == BF16
        s = self._stream
        H1, W1 = (H0 - 1) // 2 + 1, (T - 1) // 2 + 1          # conv1 output
        H, W = (H1 + 6 - 7) // 2 + 1, (W1 + 6 - 7) // 2 + 1     # conv2 output
        fmaps, geoms = [], []
        fast = bf and self.bb_kind == "basic"      # flat-layout tcgen05 path of the default net
        if self.bb_kind == "custom":
            fmaps, geoms = self._run_custom_backbone(xs, plan)
            if taps is not None:
                taps["fmaps"] = [f.float().permute(0, 3, 1, 2).contiguous() for f in fmaps]
        elif fast:
            # conv1 writes the space-to-depth flat layout [B, W2+2, H2+2, 4 parity planes x 64]; conv2 (7x7 stride 2) then is a
            # stride-1 flat conv with 49 (plane, shift) steps; backbone activations stay in the flat layout (conv_flat.cu)
            H2, W2 = (H1 + 1) // 2, (W1 + 1) // 2
            assert (H2, W2) == (H, W)
            if self.fused_stem and H0 == 32 and 32 <= T <= 1024 and "xs_bf16" in plan:
                cur = self._flat_buf(plan, "c2", B, H, W, 64)
                Hp, Wp = self._flat_geom(H, W)
                xb = plan["xs_bf16"]
                fx = plan.get("fstem_cols")
                if fx is None:
                    cols, var = self._fused_stem_border_cols(T)
                    Wo_ = W
                    lo_n = sum(1 for c_ in cols if c_ < 2)
                    hi_n = len(cols) - lo_n
                    contiguous = cols == list(range(lo_n)) + list(range(Wo_ - hi_n, Wo_))
                    fx = plan["fstem_cols"] = ((C.c_int32 * len(cols))(*cols), (C.c_int32 * len(var))(*var), len(cols), lo_n, hi_n, contiguous)
                nint = int(os.environ.get("YAD_STEM_NINT", "0"))
                if self.stem_overlap and fx[5] and fx[2] > 0:
                    # the border columns on a second stream, concurrently with the tensor-core kernel (which skips them)
                    self._pseudo(self._fork, "fork")
                    _lib.check(self.lib.yad_conv_stem_fused_fixup_bf16(xb.data_ptr(), xb.shape[2], B, H0, T, self.fstem_wvar.data_ptr(),
                                                                       self.fstem_bias.data_ptr(), fx[0], fx[1], fx[2], cur.data_ptr(), Hp, Wp,
                                                                       _SideHandle(self._side_stream().cuda_stream)), "conv_stem_fused_fixup_bf16")
                    _lib.check(self.lib.yad_conv_stem_fused_skip(xb.data_ptr(), xb.shape[2], B, H0, T, self.fstem_w.data_ptr(),
                                                                 self.fstem_bias.data_ptr(), cur.data_ptr(), Hp, Wp, nint, fx[3], fx[4], s()),
                               "conv_stem_fused_skip")
                    self._pseudo(self._join, "join")
                else:
                    _lib.check(self.lib.yad_conv_stem_fused(xb.data_ptr(), xb.shape[2], B, H0, T, self.fstem_w.data_ptr(), self.fstem_bias.data_ptr(),
                                                            cur.data_ptr(), Hp, Wp, nint, s()), "conv_stem_fused")
                    _lib.check(self.lib.yad_conv_stem_fused_fixup(xs.data_ptr(), B, H0, T, self.fstem_wvar.data_ptr(), self.fstem_bias.data_ptr(),
                                                                  fx[0], fx[1], fx[2], cur.data_ptr(), Hp, Wp, s()), "conv_stem_fused_fixup")
            else:
                cur = self._run_stem_two_convs(xs, plan, B, H0, T, H, W, H2, W2)
            s2d_in = None      # space-to-depth copy of `cur` (written by the previous layer's last convolution)
            for li, blocks in enumerate(self.stages):
                for bi, blk in enumerate(blocks):
                    c1v, c2v = blk["c1"], blk["c2"]
                    Ho, Wo = (H + 2 - 3) // c1v.sh + 1, (W + 2 - 3) // c1v.sw + 1
                    ld = _ceil(c1v.cout, 64)
                    t = self._flat_buf(plan, f"s{li}.{bi}.t", B, Ho, Wo, ld)
                    y = self._flat_buf(plan, f"s{li}.{bi}.y", B, Ho, Wo, ld)
                    idt = self._flat_buf(plan, f"s{li}.{bi}.d", B, Ho, Wo, ld) if "ds" in blk else cur
                    if c1v.sh == 1 and c1v.sw == 1:
                        self._conv_flat(c1v, cur, H, W, t)
                        if "ds" in blk:
                            self._conv_flat_s2(blk["ds"], cur, H, W, idt, Ho, Wo)
                    elif s2d_in is not None:      # stride-2 block on the space-to-depth copy: both convs in the patch-resident kernel
                        self._conv_s2d(c1v, s2d_in, H, W, t, Ho, Wo)
                        fold_ds = "ds" in blk and self.fold_ds and bi < len(blocks) - 1 and blk["ds"].cout_pad == c2v.cout_pad
                        if "ds" in blk and not fold_ds:
                            self._conv_s2d(blk["ds"], s2d_in, H, W, idt, Ho, Wo)
                        if fold_ds:               # ... and the downsample branch rides in conv2's GEMM
                            self._conv_flat_plus_ds(blk, t, s2d_in, Ho, Wo, y)
                            cur, H, W, s2d_in = y, Ho, Wo, None
                            continue
                    else:
                        dual = ("ds" in blk and self.dual_ds and blk["ds"].cout_pad == c1v.cout_pad == c1v.cout and c1v.cout % 32 == 0
                                and (blk["ds"].sh, blk["ds"].sw) == (c1v.sh, c1v.sw))
                        if dual:       # conv1 and the downsample branch of the same input in one tap-by-tap launch
                            self._conv_flat_s2(c1v, cur, H, W, t, Ho, Wo, ds=blk["ds"], out_ds=idt)
                        else:
                            self._conv_flat_s2(c1v, cur, H, W, t, Ho, Wo)
                            if "ds" in blk:
                                self._conv_flat_s2(blk["ds"], cur, H, W, idt, Ho, Wo)
                    s2d_in = None
                    nxt = self.stages[li + 1][0] if (bi == len(blocks) - 1 and li + 1 < len(self.stages)) else None
                    if (nxt is not None and self.s2d_route and (nxt["c1"].sh, nxt["c1"].sw) == (2, 2) and nxt["c1"].kh == 3 and
                            nxt["c1"].ph == 1 and c2v.cout == c2v.cout_pad == nxt["c1"].cin_pad and c2v.cout % 64 == 0 and
                            ("ds" not in nxt or ((nxt["ds"].sh, nxt["ds"].sw) == (2, 2) and nxt["ds"].kh == 1 and nxt["ds"].ph == 0))):
                        H2n, W2n = (Ho + 2 - 3) // 2 + 1, (Wo + 2 - 3) // 2 + 1
                        s2d_in = plan.get(f"s{li}.s2d")
                        if s2d_in is None:
                            Hp2, Wp2 = self._flat_geom(H2n, W2n)
                            s2d_in = plan[f"s{li}.s2d"] = torch.zeros((B, Wp2, Hp2, 4 * c2v.cout), device=self.dev, dtype=torch.bfloat16)
                        self._conv_flat_with_s2d(c2v, t, Ho, Wo, y, idt, s2d_in)
                    else:
                        self._conv_flat(c2v, t, Ho, Wo, y, res=idt)
                    cur, H, W = y, Ho, Wo
                fmaps.append(cur)
                geoms.append((H, W))
            if taps is not None:
                taps["fmaps"] = [f[:, :w_, :h_, :].float().permute(0, 3, 2, 1).contiguous() for f, (h_, w_) in zip(fmaps, geoms)]
        else:
            # dense NHWC path: fp32 parity mode, and the Bottleneck ResNet in bf16 (generic tap-by-tap tcgen05 convs)
            c1 = self._buf(plan, "c1", B, H1, W1, 64)
            if bf:
                _lib.check(self.lib.yad_conv_stem_tc(xs.data_ptr(), B, H0, T, self.stem_w_tc.data_ptr(), c1.data_ptr(), 0, 0, s()),
                           "conv_stem_tc (dense)")
            else:
                _lib.check(self.lib.yad_conv_stem(xs.data_ptr(), B, H0, T, self.stem_w.data_ptr(), c1.data_ptr(), self.dtype, s()),
                           "conv_stem")
            cur = self._buf(plan, "c2", B, H, W, 64)
            self._conv(self.conv2, c1, 0, cur, 0)
            for li, blocks in enumerate(self.stages):
                for bi, blk in enumerate(blocks):
                    if "c3" in blk:       # Bottleneck: 1x1 - 3x3 (stride) - 1x1 (+ identity / downsample) - ReLU
                        st = blk["c2"].sh
                        Ho, Wo = (H + 2 - 3) // st + 1, (W + 2 - 3) // st + 1
                        t1 = self._buf(plan, f"s{li}.{bi}.t1", B, H, W, blk["c1"].cout)
                        t2 = self._buf(plan, f"s{li}.{bi}.t2", B, Ho, Wo, blk["c2"].cout)
                        y = self._buf(plan, f"s{li}.{bi}.y", B, Ho, Wo, blk["c3"].cout)
                        self._conv(blk["c1"], cur, 0, t1, 0)
                        self._conv(blk["c2"], t1, 0, t2, 0)
                        if "ds" in blk:
                            idt = self._buf(plan, f"s{li}.{bi}.d", B, Ho, Wo, blk["c3"].cout)
                            self._conv(blk["ds"], cur, 0, idt, 0)
                        else:
                            idt = cur
                        self._conv(blk["c3"], t2, 0, y, 0, res=idt)
                        cur, H, W = y, Ho, Wo
                        continue
                    c1v, c2v = blk["c1"], blk["c2"]
                    Ho, Wo = (H + 2 - 3) // c1v.sh + 1, (W + 2 - 3) // c1v.sw + 1
                    ld = c1v.cout
                    t = self._buf(plan, f"s{li}.{bi}.t", B, Ho, Wo, ld)
                    y = self._buf(plan, f"s{li}.{bi}.y", B, Ho, Wo, ld)
                    self._conv(c1v, cur, 0, t, 0)
                    if "ds" in blk:
                        idt = self._buf(plan, f"s{li}.{bi}.d", B, Ho, Wo, ld)
                        self._conv(blk["ds"], cur, 0, idt, 0)
                    else:
                        idt = cur
                    self._conv(c2v, t, 0, y, 0, res=idt)
                    cur, H, W = y, Ho, Wo
                fmaps.append(cur)
                geoms.append((H, W))
            if taps is not None:
                taps["fmaps"] = [f.float().permute(0, 3, 1, 2).contiguous() for f in fmaps]

        # ---- neck (H = 1 after the H-mean; modules/_common.py:241-265)
        hs = [g_[0] for g_ in geoms]
        if fast and self.fused_neck and taps is None and hs[0] != hs[1] != hs[2] != hs[3]:
            fn = self._fused_neck_for(geoms, fmaps)
            if fn is not None:
                # the whole neck (H-means included) as ONE persistent kernel reading the flat backbone maps (neck_fused.cu)
                h32 = _ceil(self.n_head, 4)
                heads = [self._buf(plan, f"n{i + 2}f", B, 1, geoms[i + 1][1], h32, zero=True, dtype=torch.float32) for i in range(3)]
                fn.run(self.lib, fmaps, heads, s())
                return heads
        # the reference's chained comparison (modules/_common.py:248): the H-mean runs only when it is true; otherwise (equal
        # heights, i.e. the custom backbone) the neck stays 2-D and the heads are averaged over H at the very end (:259-261)
        pool_first = hs[0] != hs[1] != hs[2] != hs[3]
        if not pool_first and len(set(hs)) != 1:
            raise NotImplementedError(f"feature-map heights {hs}: the reference's neck would fail on these shapes as well")
        Hn = 1 if pool_first else hs[0]
        pooled = []
        for i, (f, (Hf, Wf)) in enumerate(zip(fmaps, geoms)):
            ldf = f.shape[3]
            if not pool_first or (Hf == 1 and not fast):
                pooled.append(f)
                continue
            if Hf == 1 and fast:
                # the flat layout of an H = 1 stage is already [B, Wp, 1, ld]: the mean is the identity, the neck's 1x1 convs read
                # it in place through a strided view (batch pitch Wp pixels) instead of a copy kernel
                Hpf, Wpf = self._flat_geom(Hf, Wf)
                pooled.append(f.as_strided((B, 1, Wf, ldf), (Wpf * Hpf * ldf, Wpf * Hpf * ldf, Hpf * ldf, 1)))
                continue
            pm = self._buf(plan, f"fm{i}", B, 1, Wf, ldf)
            if fast:
                Hpf, Wpf = self._flat_geom(Hf, Wf)
                isw, ish, isb = Hpf, 1, Wpf * Hpf
            else:
                isw, ish, isb = 1, Wf, Hf * Wf
            _lib.check(self.lib.yad_hmean(f.data_ptr(), self.dtype, B, Hf, Wf, ldf, ldf, isw, ish, isb, pm.data_ptr(), ldf, 0, s()), "hmean")
            pooled.append(pm)
        f1m, f2m, f3m, f4m = pooled
        W1, W2, W3, W4 = f1m.shape[2], f2m.shape[2], f3m.shape[2], f4m.shape[2]
        if not (W1 == 2 * W2 and W2 == 2 * W3 and W3 == 2 * W4):
            # the BiC blocks concatenate a x0.5, a x1 and a x2 map (modules/_common.py:179-185): the reference fails in torch.cat
            # for such lengths (e.g. 5 s or 15 s clips); the resize kernels derive their output width from the input, so an
            # un-nested pyramid would write past the cat buffers
            raise ValueError(f"feature-map widths {W1, W2, W3, W4} do not nest (each must be twice the next): the neck's "
                             "concatenations are undefined for this input length (the reference raises in torch.cat)")
        n = self.n
        # CSPSPPF -> p4, written straight into cat_n4[:, 0:128]
        a1 = self._buf(plan, "sp.a1", B, Hn, W4, 64)
        a2 = self._buf(plan, "sp.a2", B, Hn, W4, 64)
        cat5 = self._buf(plan, "sp.cat5", B, Hn, W4, 256)
        cat7 = self._buf(plan, "sp.cat7", B, Hn, W4, 128)
        c5 = self._buf(plan, "sp.c5", B, Hn, W4, 64)
        cat_n4 = self._buf(plan, "cat_n4", B, Hn, W4, 256)
        self._conv(n["sp1"], f4m, 0, a1, 0)
        self._conv(n["sp3"], a1, 0, a2, 0)
        self._conv(n["sp4"], a2, 0, cat5, 0)
        self._conv(n["sp2"], f4m, 0, cat7, 64)
        if Hn == 1:
            _lib.check(self.lib.yad_sppf_pools(cat5.data_ptr(), self.dtype, B, W4, 64, 256, 0, cat5.data_ptr(), 256, 64, s()), "sppf")
        else:
            # 2-D 5x5 cascades = W cascades (into a scratch tensor) followed by box maxima over 5 / 9 / 13 rows
            wp = self._buf(plan, "sp.wpool", B, Hn, W4, 192)
            _lib.check(self.lib.yad_sppf_pools(cat5.data_ptr(), self.dtype, B * Hn, W4, 64, 256, 0, wp.data_ptr(), 192, 0, s()), "sppf(W)")
            for k_ in range(3):
                _lib.check(self.lib.yad_maxpool_h(wp.data_ptr(), self.dtype, B, Hn, W4, 64, 192, 64 * k_, 2 * (k_ + 1),
                                                  cat5.data_ptr(), 256, 64 * (k_ + 1), s()), "sppf(H)")
        self._conv(n["sp5"], cat5, 0, c5, 0)
        self._conv(n["sp6"], c5, 0, cat7, 0)
        self._conv(n["sp7"], cat7, 0, cat_n4, 0)                       # p4
        # BiC3 -> RepBlock3_1 -> p3, written into cat_n3[:, 0:128]
        cat_b3 = self._buf(plan, "cat_b3", B, Hn, W3, 256)
        c0_3 = self._buf(plan, "b3.c0", B, Hn, W2, 64)
        b3 = self._buf(plan, "b3", B, Hn, W3, 128)
        cat_n3 = self._buf(plan, "cat_n3", B, Hn, W3, 256)
        self._conv(n["b3c1"], f3m, 0, cat_b3, 0)
        self._conv(n["b3c0"], f2m, 0, c0_3, 0)
        _lib.check(self.lib.yad_resize_w(c0_3.data_ptr(), self.dtype, B * Hn, W2, 64, 64, 0, 0, cat_b3.data_ptr(), 256, 64, s()), "pairavg")
        _lib.check(self.lib.yad_resize_w(cat_n4.data_ptr(), self.dtype, B * Hn, W4, 128, 256, 0, 1, cat_b3.data_ptr(), 256, 128, s()), "up2")
        self._conv(n["b3o"], cat_b3, 0, b3, 0)
        self._repblock(plan, "rb31", self.rep["rep_block3_1"], b3, 0, cat_n3, 0)   # p3
        # BiC2 -> RepBlock2_1 -> n2 (sm head)
        cat_b2 = self._buf(plan, "cat_b2", B, Hn, W2, 256)
        c0_2 = self._buf(plan, "b2.c0", B, Hn, W1, 64)
        b2 = self._buf(plan, "b2", B, Hn, W2, 128)
        self._conv(n["b2c1"], f2m, 0, cat_b2, 0)
        self._conv(n["b2c0"], f1m, 0, c0_2, 0)
        _lib.check(self.lib.yad_resize_w(c0_2.data_ptr(), self.dtype, B * Hn, W1, 64, 64, 0, 0, cat_b2.data_ptr(), 256, 64, s()), "pairavg")
        _lib.check(self.lib.yad_resize_w(cat_n3.data_ptr(), self.dtype, B * Hn, W3, 128, 256, 0, 1, cat_b2.data_ptr(), 256, 128, s()), "up2")
        self._conv(n["b2o"], cat_b2, 0, b2, 0)
        hl = _ceil(self.n_head, 64) if bf else self.n_head     # head channel pitch
        h32 = _ceil(self.n_head, 4)                            # fp32 head copies fed to the decoder
        n2 = self._buf(plan, "n2", B, Hn, W2, hl, zero=True)
        n3 = self._buf(plan, "n3", B, Hn, W3, hl, zero=True)
        n4 = self._buf(plan, "n4", B, Hn, W4, hl, zero=True)
        if bf:
            n2f = self._buf(plan, "n2f", B, Hn, W2, h32, zero=True, dtype=torch.float32)
            n3f = self._buf(plan, "n3f", B, Hn, W3, h32, zero=True, dtype=torch.float32)
            n4f = self._buf(plan, "n4f", B, Hn, W4, h32, zero=True, dtype=torch.float32)
        else:
            n2f, n3f, n4f = None, None, None
        self._repblock(plan, "rb21", self.rep["rep_block2_1"], b2, 0, n2, 0, out2=n2f)
        self._conv(n["ds2"], n2, 0, cat_n3, 128)
        self._repblock(plan, "rb32", self.rep["rep_block3_2"], cat_n3, 0, n3, 0, out2=n3f)
        self._conv(n["ds3"], n3, 0, cat_n4, 128)
        self._repblock(plan, "rb41", self.rep["rep_block4_1"], cat_n4, 0, n4, 0, out2=n4f)
        heads = [n2f, n3f, n4f] if bf else [n2, n3, n4]
        if Hn > 1:     # adaptive_avg_pool2d of the three heads over H (modules/_common.py:259-261), fp32
            pooled_heads = []
            for i, h in enumerate(heads):
                Wh, ldh = h.shape[2], h.shape[3]
                ph = self._buf(plan, f"head_mean{i}", B, 1, Wh, ldh, dtype=torch.float32)
                _lib.check(self.lib.yad_hmean(h.data_ptr(), F32, B, Hn, Wh, ldh, ldh, 1, Wh, Hn * Wh, ph.data_ptr(), ldh, 0, s()),
                           "hmean(heads)")
                pooled_heads.append(ph)
            heads = pooled_heads
        if taps is not None:
            taps["heads"] = [h.reshape(B, h.shape[2], h.shape[3])[..., : self.n_head].float().clone() for h in heads]
        return heads

    def _run_stem_two_convs(self, xs, plan, B, H0, T, H, W, H2, W2):
        """conv1 -> space-to-depth flat tensor -> conv2 as a stride-1 flat conv (the path the fused stem replaces; kept for input
        heights other than 32 and as the A/B reference: YAD_FUSED_STEM=0)."""
        s = self._stream
        c1 = plan.get("c1")
        if c1 is None:
            c1 = plan["c1"] = torch.zeros((B, W2 + 2, H2 + 2, 256), device=self.dev, dtype=torch.bfloat16)
        _lib.check(self.lib.yad_conv_stem_tc(xs.data_ptr(), B, H0, T, self.stem_w_tc.data_ptr(), c1.data_ptr(), H2 + 2, W2 + 2, s()),
                   "conv_stem_tc")
        cur = self._flat_buf(plan, "c2", B, H, W, 64)
        Hp, Wp = self._flat_geom(H, W)
        cv = self.conv2
        d = FlatDesc(B=B, H=H2, W=W2, Hp=H2 + 2, Wp=W2 + 2, Cin=256, ld_in=256, Cout=cv.cout, ld_out=64, co_off=0, kh=1, kw=1,
                     ph=0, pw=0, act=cv.act, ld_res=0, Hp_out=Hp, Wp_out=Wp)
        ch, dh, dw, wk = self.conv2_steps
        _lib.check(self.lib.yad_conv_flat_taps(C.byref(d), len(ch), ch, dh, dw, wk, cv.kh * cv.kw * cv.cin_pad, c1.data_ptr(),
                                               cv.w.data_ptr(), cv.cout_pad, cv.bias.data_ptr(), 0, cur.data_ptr(), s()),
                   "conv fe.conv2 (flat, space-to-depth)")
        return cur

    def _run_custom_backbone(self, xs: torch.Tensor, plan: dict):
        """CustomBackBone.forward in eval mode (modules/_backbone.py:108-116), dense NHWC; every ExtractorLayer writes its two
        branches into the channel slices of one buffer (the reference's torch.cat, :45)."""
        B, Cx, H, W = xs.shape
        s = self._stream
        x0 = self._buf(plan, "cb.x", B, H, W, Cx)
        _lib.check(self.lib.yad_nchw_to_nhwc(xs.data_ptr(), B, Cx, H, W, x0.data_ptr(), self.dtype, Cx, s()), "nchw_to_nhwc")
        pitch = (lambda c: _ceil(c, 64)) if self.dtype == BF16 else (lambda c: c)
        cur = self._buf(plan, "cb.first", B, H, W, pitch(64))
        self._conv(self.first_conv, x0, 0, cur, 0)
        fmaps, geoms = [], []
        for bi, layers in enumerate(self.cblocks):
            for li, lay in enumerate(layers):
                ca, cb, cr = lay["a"], lay["b"], lay["r"]
                Wo = (W + 2 * ca.pw - ca.kw) // ca.sw + 1
                Ho = (H + 2 * cb.ph - cb.kh) // cb.sh + 1
                t = self._buf(plan, f"cb{bi}.{li}.t", B, H, Wo, pitch(ca.cout), zero=True)
                y = self._buf(plan, f"cb{bi}.{li}.y", B, Ho, Wo, pitch(cb.cout + cr.cout), zero=True)
                self._conv(ca, cur, 0, t, 0)
                self._conv(cb, t, 0, y, 0)
                self._conv(cr, cur, 0, y, cb.cout)
                cur, H, W = y, Ho, Wo
            if bi >= 1:
                fmaps.append(cur)
                geoms.append((H, W))
        return fmaps, geoms

    def run_decode(self, heads: List[torch.Tensor], B: int, T: int, L_res: int) -> torch.Tensor:
        G = [h.shape[2] for h in heads]
        rows = sum(G) * self.A
        preds = torch.empty((B, rows, self.E), device=self.dev, dtype=torch.float32)
        hp = (C.c_void_p * 3)(*[h.data_ptr() for h in heads])
        Gs = (C.c_int32 * 3)(*G)
        lds = (C.c_int32 * 3)(*[h.shape[3] for h in heads])
        st = (C.c_int32 * 3)(*[T // g for g in G])                       # stride = spectral_size // grid_size
        anc = (C.c_float * (3 * self.A))(*self.anchors_s.tolist())
        center_scaler = T / (L_res / self.cfg["new_sample_rate"])        # _architecture.py:145
        rc = self.lib.yad_decode(hp, Gs, lds, st, 3, F32, anc, self.A, self.nc, center_scaler, float(self.cfg["sample_duration"]),
                                 B, preds.data_ptr(), self._stream())
        _lib.check(rc, "decode")
        return preds

    def _plan(self, key) -> dict:
        plans = getattr(self._tls, "plans", None)
        if plans is None:
            plans = self._tls.plans = {}
        if key not in plans:
            if len(plans) >= 8:          # bound the cached workspaces per thread
                plans.pop(next(iter(plans)))
            plans[key] = {}
        return plans[key]

    def run(self, x: torch.Tensor, taps: Optional[dict] = None, taper: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [B,1,L] f32 (or int16 PCM) on the engine's device -> preds [B, P, 3+nc] f32 (combined scales)."""
        if x.device != self.dev:
            raise RuntimeError(f"input on {x.device}, model on {self.dev}")
        B, _, L = x.shape
        plan = self._plan((B, L))
        # the first forward of a plan is recorded (every C-ABI call with its arguments), later ones replay the list with the
        # input pointer, the freshly allocated output and the current stream patched in
        fast = taps is None and x.is_contiguous() and x.dtype in (torch.float32, torch.int16)
        pkey = ("prog", x.dtype, None if taper is None else taper.data_ptr())
        if fast:
            prog = plan.get(pkey)
            if prog is not None:
                return self._replay_or_graph(prog, x)
        with torch.cuda.device(self.dev):
            if fast:
                self._tls.rec = []
            try:
                xs = self.run_frontend(x, plan, taps, taper)
                heads = self.run_cnn(xs, plan, taps)
                L_res = -(-self.rs_P * L // self.rs_O)
                preds = self.run_decode(heads, B, xs.shape[-1], L_res)
            finally:
                rec, self._tls.rec = getattr(self._tls, "rec", None), None
            if fast and rec:
                plan[pkey] = self._compile(rec, x, preds)
            return preds

    @staticmethod
    def _compile(rec, x: torch.Tensor, preds: torch.Tensor) -> dict:
        in_ptr, out_ptr = x.data_ptr(), preds.data_ptr()
        pin, pout, pst = [], [], []
        for ci, (_, args, _) in enumerate(rec):
            for ai, a in enumerate(args):
                if isinstance(a, _SideHandle):
                    continue
                if isinstance(a, C.c_void_p):
                    pst.append((ci, ai))
                elif isinstance(a, int) and not isinstance(a, bool):
                    if a == in_ptr:
                        pin.append((ci, ai))
                    elif a == out_ptr:
                        pout.append((ci, ai))
        if not pin or not pout:
            raise RuntimeError("engine: recorded call list does not reference the input / output tensors")
        return {"calls": rec, "in": pin, "out": pout, "stream": pst, "shape": tuple(preds.shape),
                "n_launch": sum(1 for _, args, _ in rec if args)}        # fork / join pseudo-calls launch nothing

    def _replay_or_graph(self, prog: dict, x: torch.Tensor) -> torch.Tensor:
        """Replay of a recorded plan; when the SAME input tensor comes back (a staging buffer, a benchmark loop) the call list is
        captured into a CUDA graph once and the whole forward becomes one graph launch.  The host then needs microseconds per
        step: on a busy machine the Python thread is descheduled for tens of milliseconds now and then, and with 57 launches to
        issue per 3.9 ms step that showed up as 6-15 ms steps in a third of the benchmark runs."""
        # graphs only while ONE thread uses this engine: a capture forbids device-wide operations (allocation, synchronize) that
        # other worker threads of the reference's 10-thread fan-out (inference.py:212-236) issue at any time
        self._threads_seen.add(threading.get_ident())
        if not self.use_graphs or len(self._threads_seen) > 1 or torch.cuda.is_current_stream_capturing():
            return self._replay(prog, x)
        # one graph per input buffer (up to 4: a ring of staging buffers / the benchmark's rotating batches)
        graphs = prog.setdefault("graphs", {})
        g = graphs.get(x.data_ptr())
        if g is not None:
            if g["x"]() is x:
                g["graph"].replay()
                _lib.launch_count += prog["n_launch"]
                return g["preds"].clone()
            graphs.pop(x.data_ptr())            # the captured input tensor is gone: its memory is somebody else's now
        for k in [k for k, v in graphs.items() if v["x"]() is None]:
            graphs.pop(k)
        seen = prog.setdefault("last", {})
        s_ = seen.get(x.data_ptr())
        if s_ is not None and s_() is x and len(graphs) < 4 and not prog.get("no_graph"):
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.dev)
            try:
                # thread_local: the reference shares one model across 10 worker threads (inference.py:212-236); their CUDA
                # calls must not be poisoned by this thread's capture
                with torch.cuda.device(self.dev), torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    preds = self._replay(prog, x)
            except Exception:  # noqa: BLE001  - capture is an optimisation: fall back to the plain replay for this plan
                prog["no_graph"] = True
                torch.cuda.synchronize(self.dev)
                return self._replay(prog, x)
            graphs[x.data_ptr()] = {"graph": graph, "preds": preds, "x": weakref.ref(x)}
            graph.replay()
            return preds.clone()
        if len(seen) >= 8:
            seen.pop(next(iter(seen)))
        seen[x.data_ptr()] = weakref.ref(x)
        return self._replay(prog, x)

    def _replay(self, prog: dict, x: torch.Tensor) -> torch.Tensor:
        calls = prog["calls"]
        with torch.cuda.device(self.dev):
            preds = torch.empty(prog["shape"], device=self.dev, dtype=torch.float32)
            st = self._stream()
            ip, op = x.data_ptr(), preds.data_ptr()
            for ci, ai in prog["in"]:
                calls[ci][1][ai] = ip
            for ci, ai in prog["out"]:
                calls[ci][1][ai] = op
            for ci, ai in prog["stream"]:
                calls[ci][1][ai] = st
            for fn, args, name in calls:
                rc = fn(*args)
                if rc:
                    _lib.check(rc, name)
            _lib.launch_count += prog["n_launch"]
        return preds
