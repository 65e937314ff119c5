"""Drop-in ``AudioDetectionNetwork`` (reference: modules/_architecture.py:10-189).

Same constructor, ``forward(x, combine_scales)``, ``inference()``, ``init_zeros_taper_window`` and
state-dict key set as the reference, so a reference checkpoint loads unchanged and the object can be
handed to the reference's ``inference.py`` / ``pipeline/_trainer.py``.  The ``nn.Module`` tree below only
*holds* parameters and buffers (names mirror modules/_common.py and modules/_backbone.py); ``forward``
never calls ``nn.Conv2d.forward`` - it runs the sm_100a kernels through :mod:`engine`.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, Optional, Tuple, Union

import torch
import torch.nn as nn
import yaml

from . import frontend_consts as fc
from .config import DEFAULT_CONFIG_PATH


def _unsupported(what: str):
    raise NotImplementedError(f"yad_b200: {what} is not built (no silent fallback); see DESIGN.md 'out of scope'")


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - holders are never called
        raise RuntimeError("parameter holder: the computation runs in yad_b200.engine, not in nn.Module.forward")


# ---------------------------------------------------------------- neck holders (modules/_common.py)
class ConvBorINorm(_NoForward):
    """conv(+bias) -> BatchNorm2d -> LeakyReLU(0.2).  ref: modules/_common.py:7-48."""

    def __init__(self, cin: int, cout: int, kernel_size, stride=1, padding=None, bias: bool = True, activation: bool = True):
        super().__init__()
        ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        if padding is None:
            padding = tuple(k // 2 for k in ks)
        self.conv = nn.Conv2d(cin, cout, kernel_size=ks, stride=stride, padding=padding, bias=bias)
        self.norm = nn.BatchNorm2d(cout)
        self.has_activation = activation


class RepVGGBlock(_NoForward):
    """ref: modules/_common.py:51-145.  Train-form: conv3x3 / conv1x1 / identity-BN branches, each branch
    ConvBorINorm keeps its own LeakyReLU.  After ``toggle_inference_mode`` only
    ``conv_reparam`` (3x3, bias) remains - a *different function* from the train-form."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.inference_mode = False
        self.conv3x3 = ConvBorINorm(cin, cout, (3, 3), padding=1, bias=False)
        self.conv1x1 = ConvBorINorm(cin, cout, (1, 1), padding=0, bias=False)
        self.identity = nn.BatchNorm2d(cout) if cin == cout else nn.Identity()

    @staticmethod
    def _merge(w: torch.Tensor, bn: nn.BatchNorm2d) -> Tuple[torch.Tensor, torch.Tensor]:
        std = torch.sqrt(bn.running_var + bn.eps)
        return (bn.weight / std).reshape(-1, 1, 1, 1) * w, ((-bn.running_mean * bn.weight) / std) + bn.bias

    @torch.no_grad()
    def reparameterize(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Linear fold of the three branches (modules/_common.py:97-133)."""
        w3, b3 = self._merge(self.conv3x3.conv.weight, self.conv3x3.norm)
        w1, b1 = self._merge(self.conv1x1.conv.weight, self.conv1x1.norm)
        w = w3 + nn.functional.pad(w1, [1, 1, 1, 1])
        b = b3 + b1
        if isinstance(self.identity, nn.BatchNorm2d):
            eye = torch.eye(self.in_channels, device=w.device, dtype=w.dtype).reshape(self.in_channels, self.in_channels, 1, 1)
            wi, bi = self._merge(eye, self.identity)
            w = w + nn.functional.pad(wi, [1, 1, 1, 1])
            b = b + bi
        return w, b

    def toggle_inference_mode(self):
        w, b = self.reparameterize()
        self.conv_reparam = nn.Conv2d(self.in_channels, self.out_channels, kernel_size=(3, 3), stride=1, padding=1)
        self.conv_reparam.weight.data = w
        self.conv_reparam.bias.data = b
        del self.conv3x3, self.conv1x1, self.identity   # a second call raises AttributeError, like the reference
        self.inference_mode = True


class RepBlock(_NoForward):
    def __init__(self, cin: int, cout: int, n: int = 2):
        super().__init__()
        self.conv1 = RepVGGBlock(cin, cout)
        self.blocks = nn.Sequential(*[RepVGGBlock(cout, cout) for _ in range(n - 1)]) if n > 1 else nn.Identity()


class BiCModule(_NoForward):
    def __init__(self, c1_in: int, c0_in: int, p2_in: int, cout: int, e: float = 0.5):
        super().__init__()
        ch = int(cout * e)
        self.conv_c1 = ConvBorINorm(c1_in, ch, 1)
        self.conv_c0 = ConvBorINorm(c0_in, ch, 1)
        self.conv_out = ConvBorINorm(ch + ch + p2_in, cout, 1)


class CSPSPPFModule(_NoForward):
    def __init__(self, cin: int, cout: int, e: float = 0.5):
        super().__init__()
        ch = int(cout * e)
        self.conv_1_3_4 = nn.Sequential(ConvBorINorm(cin, ch, 1), ConvBorINorm(ch, ch, 3), ConvBorINorm(ch, ch, 1))
        self.conv2 = ConvBorINorm(cin, ch, 1)
        self.conv5 = ConvBorINorm(ch * 4, ch, 1)
        self.conv6 = ConvBorINorm(ch, ch, 3)
        self.conv7 = ConvBorINorm(ch * 2, cout, 1)


class MultiScaleFmapModule(_NoForward):
    """RepBi-PAN neck + heads.  ref: modules/_common.py:218-265."""

    def __init__(self, c1: int, c2: int, c3: int, c4: int, out_channels: int):
        super().__init__()
        ch = 128
        self.cspsppf = CSPSPPFModule(c4, ch)
        self.bic2 = BiCModule(c2, c1, ch, ch)
        self.bic3 = BiCModule(c3, c2, ch, ch)
        self.rep_block2_1 = RepBlock(ch, out_channels)
        self.rep_block3_1 = RepBlock(ch, ch)
        self.rep_block3_2 = RepBlock(ch * 2, out_channels)
        self.rep_block4_1 = RepBlock(ch * 2, out_channels)
        self.conv2_downsample = ConvBorINorm(out_channels, ch, 3, stride=(1, 2))
        self.conv3_downsample = ConvBorINorm(out_channels, ch, 3, stride=(1, 2))


# ---------------------------------------------------------------- backbone holders (modules/_backbone.py:119-152)
class BasicBlock(_NoForward):
    """torchvision BasicBlock parameter layout (conv1, bn1, conv2, bn2, downsample.{0,1})."""

    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.stride = stride
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))
        else:
            self.downsample = None


class Bottleneck(_NoForward):
    """torchvision Bottleneck parameter layout (conv1 1x1, conv2 3x3 carrying the stride, conv3 1x1 to 4 x planes,
    downsample.{0,1}).  ref: [tv] models/resnet.py:108-163; selected by ``resnet_config.block: Bottleneck``
    (modules/_backbone.py:128-138)."""
    expansion = 4

    def __init__(self, cin: int, planes: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, 1, 0, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, 1, 0, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.stride = stride
        if stride != 1 or cin != planes * 4:
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))
        else:
            self.downsample = None


class ResNetBackBone(_NoForward):
    def __init__(self, in_channels: int, dropout: float = 0.0, block: str = "BasicBlock",
                 block_layers: Optional[Iterable[int]] = None):
        super().__init__()
        name = block if isinstance(block, str) else getattr(block, "__name__", None)
        if name not in ("BasicBlock", "Bottleneck"):
            _unsupported(f"resnet_config.block={block!r} (BasicBlock and Bottleneck are the torchvision blocks)")
        self.block_name = name
        exp = 1 if name == "BasicBlock" else 4
        layers = list(block_layers or [3, 4, 6, 3])
        self.in_channels = in_channels
        self.dropout_p = dropout
        self.conv1 = nn.Conv2d(in_channels, 64, (7, 7), (2, 2), (3, 3), bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, (n, planes) in enumerate(zip(layers, (64, 128, 256, 512))):
            blocks = []
            for bi in range(n):
                stride = 2 if (li > 0 and bi == 0) else 1
                blocks.append(BasicBlock(cin, planes, stride) if exp == 1 else Bottleneck(cin, planes, stride))
                cin = planes * exp
            setattr(self, f"layer{li + 1}", nn.Sequential(*blocks))
        self.conv2 = nn.Conv2d(64, 64, (7, 7), (2, 2), (3, 3), bias=False)
        self.fmap1_ch, self.fmap2_ch, self.fmap3_ch, self.fmap4_ch = 64 * exp, 128 * exp, 256 * exp, 512 * exp


class ExtractorLayer(_NoForward):
    """ref: modules/_backbone.py:8-46.  ``_layer`` = conv(3x7, stride (1, sw)) - BN - LeakyReLU(0.2) - conv(3x7, stride (sh, 1)) -
    BN - Dropout (no activation after the second BatchNorm); ``_res_layer`` = biased 1x1 conv with stride (sh, sw) (the
    reference's ``if not (h_stride or w_stride)`` is never true, so it is never an Identity); output = cat(_layer, _res_layer)."""

    def __init__(self, cin: int, cout: int, dropout: float = 0.0, halve_w: bool = False, halve_h: bool = False):
        super().__init__()
        if cout % 2 == 0:
            out = res_out = cout // 2
        else:
            res_out = cout // 2
            out = cout - res_out
        sw, sh = (2 if halve_w else 1), (2 if halve_h else 1)
        self._layer = nn.Sequential(
            nn.Conv2d(cin, 32, kernel_size=(3, 7), stride=(1, sw), padding=(1, 3)),
            nn.BatchNorm2d(32),
            nn.LeakyReLU(0.2),
            nn.Conv2d(32, out, kernel_size=(3, 7), stride=(sh, 1), padding=(1, 3)),
            nn.BatchNorm2d(out),
            nn.Dropout(dropout),
        )
        self._res_layer = nn.Conv2d(cin, res_out, kernel_size=(1, 1), stride=(sh, sw))


class ExtractorBlock(_NoForward):
    """ref: modules/_backbone.py:49-79: ``num_layers`` ExtractorLayers, widths 64, 128, ... and the last one = out_channels with
    the width halved."""

    def __init__(self, cin: int, cout: int, num_layers: int, dropout: float = 0.0):
        super().__init__()
        width, md = 64, {}
        for i in range(num_layers):
            last = i + 1 == num_layers
            if last:
                width = cout
            md[f"layer{i}"] = ExtractorLayer(cin, width, dropout=dropout, halve_h=False, halve_w=last)
            cin = width
            width *= 2
        self.module_dict = nn.ModuleDict(md)


class CustomBackBone(_NoForward):
    """ref: modules/_backbone.py:82-116 (``backbone: custom``): the height stays 32 through the whole backbone, so the neck runs
    in 2-D and the heads are averaged over H at the very end (modules/_common.py:248,259-261)."""

    def __init__(self, in_channels: int, dropout: float = 0.0, block_layers: Optional[Iterable[int]] = None):
        super().__init__()
        bl = list(block_layers or [3, 4, 6, 3])
        if len(bl) != 4:
            raise ValueError("block config must be a list of length = 4")
        self.in_channels = in_channels
        self.first_conv = nn.Sequential(nn.Conv2d(in_channels, 64, kernel_size=(7, 7), stride=1, padding=3),
                                        nn.BatchNorm2d(64), nn.LeakyReLU(0.2))
        self.entry_block = ExtractorBlock(64, 64, 2, dropout=dropout)
        self.block1 = ExtractorBlock(64, 128, bl[0], dropout=dropout)
        self.block2 = ExtractorBlock(128, 256, bl[1], dropout=dropout)
        self.block3 = ExtractorBlock(256, 512, bl[2], dropout=dropout)
        self.block4 = ExtractorBlock(512, 1024, bl[3], dropout=dropout)
        self.fmap1_ch, self.fmap2_ch, self.fmap3_ch, self.fmap4_ch = 128, 256, 512, 1024


class _Buf(nn.Module):
    """Holds the registered buffers of a torchaudio transform under the reference's names."""

    def __init__(self, **bufs):
        super().__init__()
        for k, v in bufs.items():
            if isinstance(v, nn.Module):
                self.add_module(k, v)
            else:
                self.register_buffer(k, v)


# ---------------------------------------------------------------- the model
class AudioDetectionNetwork(nn.Module):
    def __init__(self, num_classes: int, config: Union[str, Dict[str, Any]] = DEFAULT_CONFIG_PATH,
                 compute_dtype: str = "bf16", train_dtype: str = "tf32"):
        super().__init__()
        self.train_graphs = False           # replay the train step from CUDA graphs (needs stable .grad buffers, e.g. FusedAdamEMA)
        self.train_dtype = train_dtype      # convolutions of train() mode: "tf32" (tcgen05, cuDNN's default for fp32 training) | "f32"
        if isinstance(config, str):
            with open(config, "r") as f:
                self.config = yaml.safe_load(f)
        elif isinstance(config, dict):
            self.config = config
        else:
            raise ValueError(f"config is expected to be str or dict type got {type(config)}")
        cfg = self.config
        self.num_classes = num_classes
        self.out_channels = cfg["num_anchors"] * (3 + num_classes)
        self.compute_dtype = compute_dtype

        mc = cfg["melspectrogram_config"]
        mk = cfg["mfcc_config"]["melkwargs"]
        for name, c in (("melspectrogram_config", mc), ("mfcc_config.melkwargs", mk)):
            if not (c["n_fft"] == 1000 and c["hop_length"] == 1000 and c["n_mels"] == 32 and c["center"] is False
                    and c["win_length"] in (None, 1000) and c["power"] == 2):
                _unsupported(f"{name}={c} (kernels are built for n_fft=hop=1000, 32 mels, center=False, power=2)")
        if mc != mk:
            _unsupported("different mel settings for MelSpectrogram and MFCC (the mel plane is computed once)")
        if cfg["mfcc_config"]["n_mfcc"] != 32:
            _unsupported(f"n_mfcc={cfg['mfcc_config']['n_mfcc']} (kernel is built for 32)")
        sr = cfg["new_sample_rate"]
        kern, self._rs_width, self._rs_orig, self._rs_new = fc.sinc_resample_kernel(cfg["sample_rate"], sr)
        fb = fc.mel_filterbank(501, 0.0, float(sr // 2), 32, sr, mc["norm"], mc["mel_scale"])
        win = torch.hann_window(1000, periodic=True)
        self.resampler = _Buf(kernel=kern)
        self.melspectogram_tfmr = _Buf(spectrogram=_Buf(window=win), mel_scale=_Buf(fb=fb))
        self.mfcc_tfmr = _Buf(dct_mat=fc.dct_ortho(32, 32),
                              MelSpectrogram=_Buf(spectrogram=_Buf(window=win.clone()), mel_scale=_Buf(fb=fb.clone())))
        self.register_buffer("taper_window", torch.empty(0), persistent=True)

        dur = cfg["sample_duration"]
        ta = cfg["train_anchors"]
        self.sm_anchors = nn.Parameter(torch.FloatTensor(cfg["anchors"]["sm"]) / dur, requires_grad=ta)
        self.md_anchors = nn.Parameter(torch.FloatTensor(cfg["anchors"]["md"]) / dur, requires_grad=ta)
        self.lg_anchors = nn.Parameter(torch.FloatTensor(cfg["anchors"]["lg"]) / dur, requires_grad=ta)

        if cfg["backbone"] == "custom":         # modules/_architecture.py:46-58
            self.feature_extractor = CustomBackBone(2, dropout=cfg["dropout"], block_layers=cfg["block_layers"])
        elif cfg["backbone"] == "resnet":
            self.feature_extractor = ResNetBackBone(2, dropout=cfg["dropout"], block_layers=cfg["block_layers"],
                                                    **cfg["resnet_config"])
        else:
            raise ValueError(f"invalid backbone {cfg['backbone']}")
        fe = self.feature_extractor
        self.multiscale_module = MultiScaleFmapModule(fe.fmap1_ch, fe.fmap2_ch, fe.fmap3_ch, fe.fmap4_ch,
                                                      out_channels=self.out_channels)
        self.apply(self.xavier_init_weights)
        self._engine_cache: Dict[Any, Any] = {}

    # ---- reference API -------------------------------------------------------------------------
    def xavier_init_weights(self, m: nn.Module):
        if isinstance(m, nn.Conv2d):
            nn.init.xavier_uniform_(m.weight)
            if torch.is_tensor(m.bias):
                m.bias.data.fill_(0.01)

    def init_zeros_taper_window(self, taper_window: torch.Tensor):
        self.taper_window = torch.zeros_like(taper_window)

    def inference(self):
        """Deploy form: fold every RepVGG block (modules/_architecture.py:171-180).  One-shot."""
        self.eval()
        for m in list(self.modules()):
            if isinstance(m, RepVGGBlock):
                m.toggle_inference_mode()
        self._engine_cache.clear()

    @staticmethod
    def scale_input(x: torch.Tensor, e: float = 1e-5) -> torch.Tensor:
        """Kept for API parity (modules/_architecture.py:182-189); the kernel path fuses this."""
        mu = x.mean(dim=(-2, -1))[:, :, None, None]
        std = x.std(dim=(-2, -1))[:, :, None, None]
        return (x - mu) / (std + e)

    def invalidate_engines(self):
        """Drop the packed-weight engines (and their CUDA-graph plans) so the next eval() forward re-packs the live parameters.
        Staleness is detected through ``tensor._version`` and ``_lib.param_epoch`` (bumped by FusedAdamEMA.step,
        EMAParamsSmoothener.update, load_state_dict, load_checkpoint); writes through ``.data`` (``p.data.mul_()``,
        ``conv.weight.data = w``) bypass both - call this after such a write."""
        from . import _lib
        _lib.param_epoch += 1
        self._engine_cache.clear()

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        from . import _lib
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        _lib.param_epoch += 1
        return out

    # ---- forward -------------------------------------------------------------------------------
    def _engine(self, frontend_only: bool = False):
        from .engine import InferenceEngine
        dev = self.sm_anchors.device
        if frontend_only:      # train mode: only the (parameter-free) frontend is packed; buffers decide staleness
            ver = sum(b._version for b in self.buffers() if b.dtype.is_floating_point and b.ndim > 1)
            key = (str(dev), "frontend")
        else:
            from . import _lib
            ver = (sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers()), _lib.param_epoch)
            key = (str(dev), self.compute_dtype)
        ent = self._engine_cache.get(key)
        if ent is None or ent[0] != ver:
            ent = (ver, InferenceEngine(self, dev, self.compute_dtype, frontend_only=frontend_only))
            self._engine_cache[key] = ent
        return ent[1]

    def _train_engine(self):
        from .train_engine import TrainEngine
        fe = self.feature_extractor
        if not isinstance(fe, ResNetBackBone):
            _unsupported("train() mode (autograd) of the custom backbone (its eval() forward is built)")
        dev = self.sm_anchors.device
        key = (str(dev), "train", self.train_dtype)
        eng = self._engine_cache.get(key)
        if eng is None:
            eng = self._engine_cache[key] = TrainEngine(self, dev, self.train_dtype)
        return eng

    def _taper(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        """``taper_input: true`` (modules/_architecture.py:87-94): the clip-long window over the RESAMPLED signal, created on first
        use with the resampled length of that input and kept as the registered buffer ``taper_window`` (it is in the state dict)."""
        if not self.config["taper_input"]:
            return None
        L_res = -(-self._rs_new * x.shape[-1] // self._rs_orig)
        if self.taper_window.numel() == 0:
            self.taper_window = getattr(torch, f"{self.config['taper_window']}_window")(L_res, periodic=False, device=x.device)
        if self.taper_window.numel() != L_res:
            raise RuntimeError(f"The size of tensor a ({L_res}) must match the size of tensor b ({self.taper_window.numel()}) at "
                               "non-singleton dimension 2 (taper_window was built for another clip length, as in the reference)")
        return self.taper_window.to(x.device, torch.float32).contiguous()

    def _forward_train(self, x: torch.Tensor):
        """train() mode (pipeline/_trainer.py:98-104): batch-statistics BatchNorm, dropout, differentiable w.r.t. every
        parameter.  fp32 throughout, like the reference's training."""
        from .train_engine import run_train_forward, run_train_forward_graphed
        fe = self._engine(frontend_only=True)
        B, _, L = x.shape
        if self.train_graphs and torch.is_grad_enabled():
            with torch.cuda.device(fe.dev):
                out = run_train_forward_graphed(self, self._train_engine(), fe, x.contiguous().float())
            if out is not None:
                return out
        with torch.no_grad(), torch.cuda.device(fe.dev):
            xs = fe.run_frontend(x, {}, None, self._taper(x))
        L_res = -(-fe.rs_P * L // fe.rs_O)
        with torch.cuda.device(fe.dev):
            return run_train_forward(self, self._train_engine(), xs, xs.shape[-1], L_res)

    def forward(self, x: torch.Tensor, combine_scales: bool = False, taps: Optional[Dict[str, torch.Tensor]] = None):
        if not x.is_cuda:
            raise RuntimeError("yad_b200.AudioDetectionNetwork runs on sm_100a only; move the input and the model to "
                               "a CUDA device (there is no CPU fallback)")
        if x.ndim != 3 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape [N, 1, n_time], got {tuple(x.shape)}")
        B, E = x.shape[0], self.num_classes + 3
        if self.training:
            sm, md, lg = self._forward_train(x)
            if combine_scales:
                return torch.cat([p.reshape(B, -1, E) for p in (sm, md, lg)], dim=1)
            return sm, md, lg
        preds = self._engine().run(x, taps=taps, taper=self._taper(x))
        if combine_scales:
            return preds
        A = self.config["num_anchors"]
        out, r0 = [], 0
        for G in self._engine().grids(x.shape[-1]):
            out.append(preds[:, r0:r0 + G * A].reshape(B, G, A, E))
            r0 += G * A
        return tuple(out)

    def __deepcopy__(self, memo):
        # engines hold device workspaces; a copy (EMA shadow model) rebuilds its own lazily
        cache, self._engine_cache = self._engine_cache, {}
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            import copy as _copy
            for k, v in self.__dict__.items():
                setattr(new, k, _copy.deepcopy(v, memo))
        finally:
            self._engine_cache = cache
        return new
