"""CPU oracle: a restatement of the reference's audio-detection hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker.  The product path
(``yad_b200``) never imports anything from ``oracle/`` and has no CPU fallback.

Everything here is plain ``torch`` fp32 on the CPU (``F.conv1d``, ``rfft``,
``F.conv2d``, ``F.batch_norm`` ...) or numpy/pure-Python loops for the integer
parts (NMS keep-sets, anchor-matching indices).  The reference itself is Python
and its arithmetic lives in three unpinned third-party wheels that are not
vendored under /root/reference (``requirements.txt:1-6`` lists none of them);
the de-facto pins are the versions of this image:

  * torchaudio 2.11.0  - Resample / MelSpectrogram / MFCC / AmplitudeToDB
  * torchvision 0.26.0 - ResNet/BasicBlock, ops.batched_nms -> torchvision::nms
  * torch 2.11.0       - everything else

Parity pinning: the reference ships **no tests and no golden vectors**
(SURVEY.md section 4), so the oracle is pinned against outputs of the live
reference imported in the dev container: ``tests/golden/make_golden.py``
imports /root/reference, runs it on seeded inputs and commits small fixtures;
``tests/test_oracle_golden.py`` checks every function here against them
(bit-exact for indices / keep-sets / constants, stated fp32 tolerance for
floats).

Each function cites the reference file:line it follows ("ref:" = relative to
/root/reference, "[ta]" = torchaudio, "[tv]" = torchvision site-packages).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------
# default configuration (ref: config/config.yaml:1-98), restated verbatim
# --------------------------------------------------------------------------
DEFAULT_CONFIG: Dict = {
    "anchors": {
        "lg": [43.19310559006211, 50.99557251908398, 59.81746359223327],
        "md": [19.551036269430053, 27.203208722741433, 35.17562231759656],
        "sm": [2.650371318822014, 7.44449691991786, 12.867792792792798],
    },
    "backbone": "resnet",
    "block_layers": [2, 2, 2, 2],
    "resnet_config": {"block": "BasicBlock"},
    "dropout": 0.4,
    "melspectrogram_config": {
        "center": False, "hop_length": 1000, "mel_scale": "htk", "n_fft": 1000, "n_mels": 32,
        "norm": "slaney", "pad_mode": "reflect", "power": 2, "win_length": None,
    },
    "mfcc_config": {
        "melkwargs": {
            "center": False, "hop_length": 1000, "mel_scale": "htk", "n_fft": 1000, "n_mels": 32,
            "norm": "slaney", "pad_mode": "reflect", "power": 2, "win_length": None,
        },
        "n_mfcc": 32,
    },
    "num_anchors": 3,
    "train_anchors": True,
    "sample_duration": 60,
    "sample_rate": 22050,
    "new_sample_rate": 16000,
    "scale_input": True,
    "taper_input": False,
    "taper_window": "hann",
    "audio_extension": "wav",
}


# --------------------------------------------------------------------------
# frontend constants
# --------------------------------------------------------------------------
def resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6,
                    rolloff: float = 0.99) -> Tuple[Tensor, int, int, int]:
    """Hann-windowed sinc polyphase bank.  [ta] functional/functional.py:1305-1400.

    Returns (kernel [new/g, 1, 2*width + orig/g] f32, width, orig/g, new/g)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = torch.arange(-width, width + o, dtype=torch.float64)[None, None] / o
    # dtype=None in the reference: the phase term is int64 / int -> *float32*, then promoted (1378-1379)
    t = torch.arange(0, -n, -1, dtype=None)[:, None, None] / n + idx
    t = t * base
    t = t.clamp(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / o
    k = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    k = k * window * scale
    return k.to(torch.float32), width, o, n


def hann_window(n: int) -> Tensor:
    """Periodic Hann ([ta] transforms/_transforms.py:96 -> torch.hann_window)."""
    return torch.hann_window(n, periodic=True, dtype=torch.float32)


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> Tensor:
    """HTK mel points, slaney area norm.  [ta] functional/functional.py:424-442,470-517,518-587."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return fb * enorm.unsqueeze(0)


def dct_matrix(n_mfcc: int, n_mels: int) -> Tensor:
    """DCT-II, ortho.  [ta] functional/functional.py:636-668.  Returns [n_mels, n_mfcc]."""
    n = torch.arange(float(n_mels))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n_mels))
    return dct.t().contiguous()


def frontend_constants(config: Dict = DEFAULT_CONFIG) -> Dict[str, Tensor]:
    """All registered buffers of the reference frontend (SURVEY App. F names)."""
    mc = config["melspectrogram_config"]
    sr = config["new_sample_rate"]
    n_fft = mc["n_fft"]
    k, width, o, n = resample_kernel(config["sample_rate"], sr)
    fb = mel_filterbank(n_fft // 2 + 1, 0.0, float(sr // 2), mc["n_mels"], sr)
    return {
        "resampler.kernel": k,
        "melspectogram_tfmr.spectrogram.window": hann_window(n_fft),
        "melspectogram_tfmr.mel_scale.fb": fb,
        "mfcc_tfmr.dct_mat": dct_matrix(config["mfcc_config"]["n_mfcc"], mc["n_mels"]),
        "mfcc_tfmr.MelSpectrogram.spectrogram.window": hann_window(n_fft),
        "mfcc_tfmr.MelSpectrogram.mel_scale.fb": fb.clone(),
    }


# --------------------------------------------------------------------------
# frontend (ref: modules/_architecture.py:84-108,182-189)
# --------------------------------------------------------------------------
def resample(x: Tensor, kernel: Tensor, orig_freq: int, new_freq: int) -> Tensor:
    """[ta] functional/functional.py:1405-1431.  x [B,1,L] -> [B,1,ceil(new*L/orig)]."""
    g = math.gcd(orig_freq, new_freq)
    o, n = orig_freq // g, new_freq // g
    width = (kernel.shape[-1] - o) // 2
    B, C, L = x.shape
    w = x.reshape(-1, L)
    w = F.pad(w, (width, width + o))
    r = F.conv1d(w[:, None], kernel, stride=o)
    r = r.transpose(1, 2).reshape(w.shape[0], -1)
    target = int(math.ceil(n * L / o))
    return r[..., :target].reshape(B, C, target)


def amplitude_to_db(x: Tensor, top_db: float = 80.0) -> Tensor:
    """[ta] functional/functional.py:390-403 with multiplier 10, amin 1e-10, ref 1.0.

    x is [B, C, H, W]; the floor is per clip over (C, H, W)."""
    x_db = 10.0 * torch.log10(torch.clamp(x, min=1e-10))
    x_db = x_db - 10.0 * math.log10(max(1e-10, 1.0))
    floor = (x_db.amax(dim=(-3, -2, -1)) - top_db).view(-1, 1, 1, 1)
    return torch.max(x_db, floor)


def scale_input(x: Tensor, e: float = 1e-5) -> Tensor:
    """ref: modules/_architecture.py:182-189 (unbiased std over H,W per clip+channel)."""
    mu = x.mean(dim=(-2, -1))[:, :, None, None]
    std = x.std(dim=(-2, -1))[:, :, None, None]
    return (x - mu) / (std + e)


def frontend(x: Tensor, consts: Dict[str, Tensor], config: Dict = DEFAULT_CONFIG) -> Dict[str, Tensor]:
    """PCM [B,1,L] f32 -> every intermediate plane of the reference frontend.

    ref: modules/_architecture.py:84-108.  Non-overlapping frames (n_fft = hop,
    center False) so torch.stft reduces to reshape * window -> rfft."""
    mc = config["melspectrogram_config"]
    n_fft, hop = mc["n_fft"], mc["hop_length"]
    assert not mc["center"] and (mc["win_length"] in (None, n_fft))
    r = resample(x, consts["resampler.kernel"], config["sample_rate"], config["new_sample_rate"])
    if config["taper_input"]:
        tw = getattr(torch, f"{config['taper_window']}_window")(r.shape[-1], periodic=False)
        r = r * tw[None, None, :]
    B = r.shape[0]
    sig = r[:, 0]
    n_frames = 1 + (sig.shape[-1] - n_fft) // hop
    if hop == n_fft:
        fr = sig[:, : n_frames * n_fft].reshape(B, n_frames, n_fft)
    else:
        fr = sig.unfold(-1, n_fft, hop)
    fr = fr * consts["melspectogram_tfmr.spectrogram.window"]
    spec = torch.fft.rfft(fr)                                   # [B, T, n_fft/2+1]
    power = spec.abs().pow(2.0)                                 # [ta] functional.py:144
    mel = torch.matmul(power, consts["melspectogram_tfmr.mel_scale.fb"]).transpose(1, 2)[:, None]  # [B,1,M,T]
    meldb_for_mfcc = amplitude_to_db(mel)                       # [ta] transforms/_transforms.py:709-716
    mfcc = torch.matmul(meldb_for_mfcc.transpose(-1, -2), consts["mfcc_tfmr.dct_mat"]).transpose(-1, -2)
    meldb = amplitude_to_db(mel)                                # ref: _architecture.py:100
    mfdb = amplitude_to_db(mfcc)                                # ref: _architecture.py:101 (second dB, Q5)
    out = {"resampled": r, "mel": mel, "meldb": meldb, "mfcc": mfcc, "mfdb": mfdb}
    if config["scale_input"]:
        meldb, mfdb = scale_input(meldb), scale_input(mfdb)
    out["x_spectral"] = torch.cat((meldb, mfdb), dim=1)
    return out


# --------------------------------------------------------------------------
# CNN (ref: modules/_backbone.py:119-152, modules/_common.py, [tv] models/resnet.py:59-105)
# --------------------------------------------------------------------------
_BN_TRAINING = [False]     # set by forward_train: BatchNorm2d in train() mode (batch statistics, running stats updated in place)


def _bn(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], _BN_TRAINING[0], 0.1, 1e-5)


def _cbl(sd, p, x, stride=1, act=True):
    """ConvBorINorm: conv(+bias) -> BN -> LeakyReLU(0.2).  ref: modules/_common.py:7-48."""
    w = sd[p + ".conv.weight"]
    b = sd.get(p + ".conv.bias")
    pad = (w.shape[2] // 2, w.shape[3] // 2)
    x = F.conv2d(x, w, b, stride=stride, padding=pad)
    x = _bn(sd, p + ".norm", x)
    return F.leaky_relu(x, 0.2) if act else x


def _repvgg(sd, p, x):
    """RepVGGBlock, both forms.  ref: modules/_common.py:86-95 (note Q1: every branch
    keeps its own LeakyReLU in train-form; deploy form is lrelu(conv_reparam(x)))."""
    if (p + ".conv_reparam.weight") in sd:
        return F.leaky_relu(F.conv2d(x, sd[p + ".conv_reparam.weight"], sd[p + ".conv_reparam.bias"],
                                     stride=1, padding=1), 0.2)
    out = F.leaky_relu(_bn(sd, p + ".conv3x3.norm", F.conv2d(x, sd[p + ".conv3x3.conv.weight"], None, 1, 1)), 0.2) \
        + F.leaky_relu(_bn(sd, p + ".conv1x1.norm", F.conv2d(x, sd[p + ".conv1x1.conv.weight"], None, 1, 0)), 0.2)
    if (p + ".identity.weight") in sd:
        out = out + _bn(sd, p + ".identity", x)
    return F.leaky_relu(out, 0.2)


def _repblock(sd, p, x):
    x = _repvgg(sd, p + ".conv1", x)
    i = 0
    while any(k.startswith(f"{p}.blocks.{i}.") for k in sd):
        x = _repvgg(sd, f"{p}.blocks.{i}", x)
        i += 1
    return x


def _basic_block(sd, p, x, stride):
    """[tv] models/resnet.py:89-105."""
    idt = x
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)))
    out = _bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1))
    if (p + ".downsample.0.weight") in sd:
        idt = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0))
    return F.relu(out + idt)


def _bottleneck(sd, p, x, stride):
    """[tv] models/resnet.py:143-163 (resnet_config.block = Bottleneck, ref: modules/_backbone.py:128-138): 1x1 - 3x3 (stride) -
    1x1 (x4), residual add, ReLU."""
    idt = x
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, 1, 0)))
    out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], None, stride, 1)))
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"], None, 1, 0))
    if (p + ".downsample.0.weight") in sd:
        idt = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0))
    return F.relu(out + idt)


def _extractor_block(sd, p, x):
    """ExtractorBlock.forward.  ref: modules/_backbone.py:49-79: only the LAST layer of a block halves the width."""
    n = 0
    while (f"{p}.module_dict.layer{n}._res_layer.weight") in sd:
        n += 1
    for i in range(n):
        q = f"{p}.module_dict.layer{i}"
        sw = 2 if i + 1 == n else 1
        x1 = F.conv2d(x, sd[q + "._layer.0.weight"], sd[q + "._layer.0.bias"], stride=(1, sw), padding=(1, 3))
        x1 = F.leaky_relu(_bn(sd, q + "._layer.1", x1), 0.2)
        x1 = F.conv2d(x1, sd[q + "._layer.3.weight"], sd[q + "._layer.3.bias"], stride=(1, 1), padding=(1, 3))
        x1 = _bn(sd, q + "._layer.4", x1)
        x2 = F.conv2d(x, sd[q + "._res_layer.weight"], sd[q + "._res_layer.bias"], stride=(1, sw))
        x = torch.cat((x1, x2), dim=1)
    return x


def custom_backbone(sd: Dict[str, Tensor], x: Tensor) -> List[Tensor]:
    """CustomBackBone.forward (eval).  ref: modules/_backbone.py:108-116."""
    p = "feature_extractor"
    x = F.conv2d(x, sd[p + ".first_conv.0.weight"], sd[p + ".first_conv.0.bias"], 1, 3)
    x = F.leaky_relu(_bn(sd, p + ".first_conv.1", x), 0.2)
    x = _extractor_block(sd, p + ".entry_block", x)
    fmaps = []
    for i in range(1, 5):
        x = _extractor_block(sd, f"{p}.block{i}", x)
        fmaps.append(x)
    return fmaps


def backbone(sd: Dict[str, Tensor], x: Tensor, block_layers: Sequence[int] = (2, 2, 2, 2)) -> List[Tensor]:
    """ResNetBackBone.forward (eval: dropout = identity).  ref: modules/_backbone.py:142-152.  The backbone family is read off
    the state dict: ``first_conv`` keys = CustomBackBone, ``conv3`` keys = torchvision Bottleneck, else BasicBlock."""
    p = "feature_extractor"
    if (p + ".first_conv.0.weight") in sd:
        return custom_backbone(sd, x)
    block = _bottleneck if (p + ".layer1.0.conv3.weight") in sd else _basic_block
    x = F.conv2d(x, sd[p + ".conv1.weight"], None, 2, 3)
    x = F.conv2d(x, sd[p + ".conv2.weight"], None, 2, 3)
    x = F.relu(_bn(sd, p + ".bn1", x))
    fmaps = []
    for li, n in enumerate(block_layers):
        for bi in range(n):
            stride = 2 if (li > 0 and bi == 0) else 1
            x = block(sd, f"{p}.layer{li + 1}.{bi}", x, stride)
        fmaps.append(x)
    return fmaps


def _bic(sd, p, c1, c0, p2):
    """BiCModule.  ref: modules/_common.py:179-185."""
    c1 = _cbl(sd, p + ".conv_c1", c1)
    c0 = F.interpolate(_cbl(sd, p + ".conv_c0", c0), scale_factor=(1, 0.5), mode="bilinear")
    p2 = F.interpolate(p2, scale_factor=(1, 2), mode="bilinear")
    return _cbl(sd, p + ".conv_out", torch.cat((c1, c0, p2), dim=1))


def _cspsppf(sd, p, x):
    """CSPSPPFModule.  ref: modules/_common.py:204-215."""
    x1 = _cbl(sd, p + ".conv_1_3_4.2", _cbl(sd, p + ".conv_1_3_4.1", _cbl(sd, p + ".conv_1_3_4.0", x)))
    y1 = _cbl(sd, p + ".conv2", x)
    p1 = F.max_pool2d(x1, 5, 1, 2)
    p2 = F.max_pool2d(p1, 5, 1, 2)
    p3 = F.max_pool2d(p2, 5, 1, 2)
    x1 = _cbl(sd, p + ".conv6", _cbl(sd, p + ".conv5", torch.cat((x1, p1, p2, p3), dim=1)))
    return _cbl(sd, p + ".conv7", torch.cat((x1, y1), dim=1))


def neck(sd: Dict[str, Tensor], fmaps: Sequence[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    """MultiScaleFmapModule.forward.  ref: modules/_common.py:241-265."""
    p = "multiscale_module"
    f1, f2, f3, f4 = fmaps
    if f1.shape[-2] != f2.shape[-2] != f3.shape[-2] != f4.shape[-2]:
        f1, f2, f3, f4 = [F.adaptive_avg_pool2d(f, (1, f.shape[-1])) for f in (f1, f2, f3, f4)]
    p4 = _cspsppf(sd, p + ".cspsppf", f4)
    p3 = _repblock(sd, p + ".rep_block3_1", _bic(sd, p + ".bic3", f3, f2, p4))
    n2 = _repblock(sd, p + ".rep_block2_1", _bic(sd, p + ".bic2", f2, f1, p3))
    n3 = _repblock(sd, p + ".rep_block3_2", torch.cat((p3, _cbl(sd, p + ".conv2_downsample", n2, stride=(1, 2))), 1))
    n4 = _repblock(sd, p + ".rep_block4_1", torch.cat((p4, _cbl(sd, p + ".conv3_downsample", n3, stride=(1, 2))), 1))
    outs = []
    for n in (n2, n3, n4):
        n = F.adaptive_avg_pool2d(n, (1, n.shape[-1]))
        outs.append(n.squeeze(2).permute(0, 2, 1))
    return tuple(outs)


def fold_repvgg(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Deploy-form state dict (AudioDetectionNetwork.inference()).

    ref: modules/_common.py:97-145, modules/_architecture.py:171-180."""
    out = dict(sd)
    blocks = sorted({k[: -len(".conv3x3.conv.weight")] for k in sd if k.endswith(".conv3x3.conv.weight")})

    def merge(w, pre):
        gamma, mu, beta = sd[pre + ".weight"], sd[pre + ".running_mean"], sd[pre + ".bias"]
        std = torch.sqrt(sd[pre + ".running_var"] + 1e-5)
        return (gamma / std).reshape(-1, 1, 1, 1) * w, ((-mu * gamma) / std) + beta

    for b in blocks:
        w3, b3 = merge(sd[b + ".conv3x3.conv.weight"], b + ".conv3x3.norm")
        w1, b1 = merge(sd[b + ".conv1x1.conv.weight"], b + ".conv1x1.norm")
        w = w3 + F.pad(w1, [1, 1, 1, 1])
        bias = b3 + b1
        if (b + ".identity.weight") in sd:
            cin = w3.shape[1]
            wi = torch.zeros((cin, cin, 1, 1))
            for i in range(cin):
                wi[i, i, 0, 0] = 1
            wI, bI = merge(wi, b + ".identity")
            w = w + F.pad(wI, [1, 1, 1, 1])
            bias = bias + bI
        for k in list(out):
            if k.startswith(b + ".conv3x3.") or k.startswith(b + ".conv1x1.") or k.startswith(b + ".identity."):
                del out[k]
        out[b + ".conv_reparam.weight"] = w
        out[b + ".conv_reparam.bias"] = bias
    return out


# --------------------------------------------------------------------------
# anchor decode (ref: modules/_architecture.py:113-156)
# --------------------------------------------------------------------------
def decode_scale(scale_pred: Tensor, anchors_s: Tensor, input_size: int, spectral_size: int,
                 num_classes: int, config: Dict = DEFAULT_CONFIG) -> Tensor:
    """get_scale_pred: [B,G,A*(3+nc)] -> [B,G,A,3+nc]; anchors_s in seconds (= param * duration)."""
    B, G, _ = scale_pred.shape
    A = anchors_s.shape[0]
    sp = scale_pred.reshape(B, G, A, -1)
    obj = sp[..., :1]
    cls = sp[..., 1:1 + num_classes]
    stride = spectral_size // G
    center_scaler = spectral_size / (input_size / config["new_sample_rate"])
    grid = torch.arange(0, G)[:, None].unsqueeze(-1)
    centers = (sp[..., -2:-1].sigmoid() * 2 - 0.5) + grid
    centers = (centers * stride) / center_scaler
    widths = (sp[..., -1:].sigmoid() * 2).pow(2) * anchors_s.unsqueeze(-1)
    centers = centers.clip(min=0, max=config["sample_duration"])
    widths = widths.clip(min=0, max=config["sample_duration"])
    return torch.cat((obj, cls, centers, widths), dim=-1)


def decode(heads: Sequence[Tensor], sd: Dict[str, Tensor], input_size: int, spectral_size: int,
           num_classes: int, config: Dict = DEFAULT_CONFIG, combine_scales: bool = True):
    """ref: modules/_architecture.py:113-130.  input_size = resampled length (16 kHz samples)."""
    dur = config["sample_duration"]
    preds = [decode_scale(h, sd[f"{n}_anchors"] * dur, input_size, spectral_size, num_classes, config)
             for h, n in zip(heads, ("sm", "md", "lg"))]
    if not combine_scales:
        return tuple(preds)
    B = heads[0].shape[0]
    preds = [p.reshape(B, -1, num_classes + 3) for p in preds]
    return torch.cat(preds, dim=1).flatten(start_dim=1, end_dim=-2)


def forward(x: Tensor, sd: Dict[str, Tensor], num_classes: int, config: Dict = DEFAULT_CONFIG,
            combine_scales: bool = True, taps: Optional[Dict] = None):
    """AudioDetectionNetwork.forward in eval() mode.  ref: modules/_architecture.py:78-130.

    Train-form if ``sd`` has conv3x3/conv1x1 keys, deploy-form if it has conv_reparam keys."""
    fe = frontend(x, sd, config)
    xs = fe["x_spectral"]
    fmaps = backbone(sd, xs, config["block_layers"])
    heads = neck(sd, fmaps)
    if taps is not None:
        taps.update(fe)
        taps["fmaps"] = fmaps
        taps["heads"] = heads
    return decode(heads, sd, fe["resampled"].shape[-1], xs.shape[-1], num_classes, config, combine_scales)


def forward_train(x: Tensor, sd: Dict[str, Tensor], num_classes: int, config: Dict = DEFAULT_CONFIG):
    """AudioDetectionNetwork.forward in train() mode with dropout = 0 (ref: modules/_architecture.py:78-130 as driven by
    pipeline/_trainer.py:98-104): every BatchNorm2d normalises with the batch statistics and updates ``sd``'s running
    statistics in place (momentum 0.1, unbiased variance).  Differentiable (torch autograd) with respect to every
    floating-point entry of ``sd`` that requires grad.  Returns the three per-scale prediction tensors."""
    if config.get("dropout", 0):
        raise ValueError("the oracle's train-mode forward is deterministic: set dropout to 0")
    _BN_TRAINING[0] = True
    try:
        return forward(x, sd, num_classes, config, combine_scales=False)
    finally:
        _BN_TRAINING[0] = False


# --------------------------------------------------------------------------
# NMS + post-processing (ref: inference.py:42-110, [tv] ops/boxes.py:20-120)
# --------------------------------------------------------------------------
def nms_greedy(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """torchvision::nms restated (binary only; behaviour pinned by probe).

    Stable descending sort; all arithmetic in fp32; suppress iff iou > (double)thr;
    a NaN iou (two zero-area boxes) never suppresses.  Returns kept indices in
    descending-score order (int64)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = ((x2 - x1) * (y2 - y1)).astype(np.float32)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = float(thr)
    zero = np.float32(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        for _i in range(n):
            i = order[_i]
            if suppressed[i]:
                continue
            keep.append(i)
            rest = order[_i + 1:]
            xx1 = np.maximum(x1[i], x1[rest])
            yy1 = np.maximum(y1[i], y1[rest])
            xx2 = np.minimum(x2[i], x2[rest])
            yy2 = np.minimum(y2[i], y2[rest])
            w = np.maximum(zero, xx2 - xx1).astype(np.float32)
            h = np.maximum(zero, yy2 - yy1).astype(np.float32)
            inter = (w * h).astype(np.float32)
            ovr = (inter / ((areas[i] + areas[rest]).astype(np.float32) - inter).astype(np.float32)).astype(np.float32)
            suppressed[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def boxes_and_confidence(outputs: Tensor, sample_duration: float = 60, _h: int = 10) -> Tuple[Tensor, Tensor]:
    """ref: inference.py:55-64.  outputs [B,P,3+nc] -> (xyxy [B,P,4], conf [B,P])."""
    cw = outputs[..., -2:]
    x1 = cw[..., :1] - (cw[..., -1:] / 2)
    x2 = cw[..., :1] + (cw[..., -1:] / 2)
    y1 = torch.zeros_like(x1)
    y2 = torch.zeros_like(x2) + _h
    coords = torch.cat([x1, y1, x2, y2], dim=-1).clip(min=0, max=sample_duration)
    objectness = outputs[..., :1].sigmoid()
    cls = F.softmax(outputs[..., 1:-2], dim=-1)
    cls = torch.gather(cls, dim=-1, index=cls.argmax(dim=-1, keepdim=True))
    return coords, (cls * objectness).squeeze(-1)


def keep_indices(outputs: Tensor, iou_threshold: float, sample_duration: float = 60, _h: int = 10) -> List[np.ndarray]:
    """Per-clip NMS keep lists (local indices, descending score).  Per-clip NMS on
    un-offset coordinates == torchvision.batched_nms whenever it picks the vanilla
    per-index loop, and at B = 1."""
    if outputs.ndim != 3:
        outputs = outputs.unsqueeze(0)
    coords, conf = boxes_and_confidence(outputs, sample_duration, _h)
    return [nms_greedy(coords[b].numpy(), conf[b].numpy(), iou_threshold) for b in range(outputs.shape[0])]


def process_model_outputs(outputs: Tensor, iou_threshold: float = 0.05, conf_threshold: float = 0.5,
                          sample_duration: float = 60, return_start_end: bool = True, _h: int = 10):
    """ref: inference.py:42-110.  Returns (segments [K,5] = [conf, obj, label, start, end], batch_idxs [K] i64).

    Raises ValueError when nothing survives, like the reference's torch.cat([]) (Q10)."""
    if outputs.ndim != 3:
        outputs = outputs.unsqueeze(0)
    coords, conf = boxes_and_confidence(outputs, sample_duration, _h)
    seg_list, bidx_list = [], []
    for b in range(outputs.shape[0]):
        keep = nms_greedy(coords[b].numpy(), conf[b].numpy(), iou_threshold)
        keep = torch.from_numpy(keep)
        kc = conf[b][keep]
        valid = kc > conf_threshold
        if int(valid.sum()) == 0:
            continue
        vo = torch.cat([kc[valid].unsqueeze(-1), outputs[b][keep][valid]], dim=-1)
        vo = vo[vo[:, -2].argsort()]
        seg_list.append(vo)
        bidx_list.append(torch.full((vo.shape[0],), b, dtype=torch.int64))
    if not seg_list:
        raise ValueError("no segment passed the confidence threshold (reference: torch.cat of an empty list)")
    seg = torch.cat(seg_list, dim=0)
    bidx = torch.cat(bidx_list, dim=0)
    if return_start_end:
        w = seg[..., -1].clone()
        seg[..., -2] = seg[..., -2] - (w / 2)
        seg[..., -1] = seg[..., -2] + w
        seg[..., -2:] = seg[..., -2:].clip(min=0, max=sample_duration)
    labels = seg[..., 2:-2].argmax(dim=-1, keepdim=True)
    seg = torch.cat([seg[..., :2], labels, seg[..., -2:]], dim=-1)
    return seg, bidx


def evaluate_waveform(wav: Tensor, sd: Dict[str, Tensor], num_classes: int, sample_rate: int, sample_duration: float,
                      batch_size: int, idx2class_map: Dict[int, str], iou_threshold: float = 0.1, conf_threshold: float = 0.65,
                      config: Dict = DEFAULT_CONFIG):
    """inference.evaluate_audio from the decoded mono waveform [n] on (ref: inference.py:126-198), incl. its clip-index quirk
    (``batch_idxs += batch_idxs_list[-1][-1]``).  ``sample_rate`` is the FILE's rate: when it differs from the model's
    ``config['sample_rate']`` every batch goes through an extra torchaudio Resample first (:152-159).
    Returns (segments in file time, batch_idxs, rle rows)."""
    from datetime import timedelta
    model_rate = int(config["sample_rate"])
    rs_kernel = resample_kernel(int(sample_rate), model_rate)[0] if int(sample_rate) != model_rate else None
    sample_size = int(sample_duration * sample_rate)
    batch_start, batch_end = 0, batch_size * sample_duration
    seg_l, idx_l = [], []
    while True:
        lo, num = int(batch_start * sample_rate), int((batch_end - batch_start) * sample_rate)
        x = wav[lo:lo + num]
        if x.shape[-1] == 0:
            break
        if x.shape[0] % sample_size != 0:
            nb = int(np.ceil(x.shape[0] / sample_size))
            x = torch.cat([x, torch.zeros((nb * sample_size - x.shape[0],), dtype=x.dtype)], dim=0)
        xb = x.reshape(-1, 1, sample_size)
        if rs_kernel is not None:
            xb = resample(xb, rs_kernel, int(sample_rate), model_rate)
        out = forward(xb, sd, num_classes, config, combine_scales=True)
        seg, bidx = process_model_outputs(out, iou_threshold, conf_threshold, sample_duration, True)
        if idx_l:
            bidx = bidx + idx_l[-1][-1]
        seg_l.append(seg)
        idx_l.append(bidx)
        batch_start = batch_end
        batch_end += batch_size * sample_duration
    seg = torch.cat(seg_l, dim=0)
    bidx = torch.cat(idx_l, dim=0)
    seg[..., -2:] = seg[..., -2:] + (bidx.unsqueeze(-1) * sample_duration)
    rows = []
    for i in range(seg.shape[0]):
        s = seg[i]
        start, end = timedelta(seconds=round(s[-2].item(), 2)), timedelta(seconds=round(s[-1].item(), 2))
        cls = idx2class_map[int(s[2].item())]
        if not rows or rows[-1]["class"] != cls:
            rows.append({"start": start, "end": end, "class": cls})
            continue
        rows[-1]["end"] = end
    return seg, bidx, rows


# --------------------------------------------------------------------------
# training-side integer work: anchor matching (ref: dataset.py:286-365)
# --------------------------------------------------------------------------
def build_target_by_scale(targets: Tensor, fmap_shape: int, anchors: Sequence[float], anchor_threshold: float = 4.0,
                          sample_duration: float = 60, edge_threshold: float = 0.5):
    """Pure-loop restatement of AudioDataset.build_target_by_scale.

    targets [T,4] = (batch_idx, cls, centre_s, dur_s) f32.  Output order: all base
    matches (anchor-major, target order), then left-neighbour copies, then right."""
    t = targets.to(torch.float32).numpy()
    anc = np.asarray(torch.tensor(anchors, dtype=torch.float32).numpy(), dtype=np.float32)
    G = fmap_shape
    base = []
    for a in range(anc.shape[0]):
        for j in range(t.shape[0]):
            r = np.float32(t[j, 3]) / anc[a]
            with np.errstate(divide="ignore"):
                m = max(r, np.float32(1.0) / r)
            if m < anchor_threshold:
                base.append((j, a))
    rows = []

    def gc_of(j):
        return np.float32(np.float32(t[j, 2] / np.float32(sample_duration)) * np.float32(G))

    for j, a in base:
        rows.append((j, a, np.float32(0.0)))
    for j, a in base:
        gc = gc_of(j)
        if (np.fmod(gc, np.float32(1)) < edge_threshold) and (gc > 1):
            rows.append((j, a, np.float32(-edge_threshold)))
    for j, a in base:
        gi = np.float32(np.float32(G) - gc_of(j))
        if (np.fmod(gi, np.float32(1)) < edge_threshold) and (gi > 1):
            rows.append((j, a, np.float32(edge_threshold)))
    bi = np.array([int(t[j, 0]) for j, a, o in rows], dtype=np.int64)
    ai = np.array([a for j, a, o in rows], dtype=np.int64)
    cl = np.array([int(t[j, 1]) for j, a, o in rows], dtype=np.int64)
    gi = np.array([min(max(int(np.float32(gc_of(j) + o)), 0), G - 1) for j, a, o in rows], dtype=np.int64)
    cw = np.array([[t[j, 2], t[j, 3]] for j, a, o in rows], dtype=np.float32).reshape(-1, 2)
    return (torch.from_numpy(bi), torch.from_numpy(gi), torch.from_numpy(ai)), torch.from_numpy(cl), torch.from_numpy(cw)


def getitem_collate(clips: Sequence[Tensor], seg_lists: Sequence[np.ndarray], sample_rate: int, sample_duration: float,
                    ignore_index: int = -100, gmins: Optional[Sequence[float]] = None) -> Tuple[Tensor, Tensor]:
    """AudioDataset.__getitem__ from the loaded waveform on + collate_fn (ref: dataset.py:123-164,276-283).  ``clips[i]`` is what
    torchaudio.load returned ([C, n] f32), ``seg_lists[i]`` the clip's (start_s, end_s, class_idx) rows."""
    audios, targets = [], []
    for i, (audio, seg) in enumerate(zip(clips, seg_lists)):
        gmin = 0.0 if gmins is None else gmins[i]
        seg = np.asarray(seg, dtype=np.float64)
        times = seg[:, :2].astype(float)
        a_start, a_end = times[0][0], times[-1][1]
        a_start, a_end = a_start - gmin, a_end - gmin
        times = times - gmin
        if audio.ndim == 1:
            audio = audio.unsqueeze(0)
        if audio.shape[0] != 1:
            audio = audio.mean(dim=0).unsqueeze(0)
        classes = torch.from_numpy(seg[:, 2].astype(np.int64))
        times[:, 1] = times[:, 1] - times[:, 0]
        times[:, 0] = times[:, 0] + (times[:, 1] / 2)
        labels = torch.cat((classes[:, None], torch.from_numpy(times).to(dtype=torch.float32)), dim=-1)
        max_n = int(sample_duration * sample_rate)
        if audio.shape[-1] < max_n:
            audio = torch.cat((audio, torch.zeros((audio.shape[0], max_n - audio.shape[-1]), dtype=audio.dtype)), dim=-1)
            pad_dur = (a_start + sample_duration) - a_end
            pad_c = a_end + (pad_dur / 2)
            labels = torch.cat((labels, torch.tensor([float(ignore_index), pad_c, pad_dur], dtype=labels.dtype).unsqueeze(0)), dim=0)
        t = torch.zeros((labels.shape[0], labels.shape[1] + 1), dtype=labels.dtype)
        t[:, 1:] = labels
        t[:, 0] = i
        audios.append(audio)
        targets.append(t)
    return torch.stack(audios, dim=0), torch.cat(targets, dim=0)


def compute_ciou(p_cw: Tensor, t_cw: Tensor, e: float = 1e-8, _h: float = 10.0) -> Tensor:
    """ref: modules/_loss.py:193-228 (1-D segments dressed as boxes of height 10)."""
    pc, pw = p_cw[..., :1], p_cw[..., -1:]
    tc, tw = t_cw[..., :1], t_cw[..., -1:]
    ph = torch.ones_like(pw) * _h
    th = torch.ones_like(tw) * _h
    px1, px2 = pc - pw / 2, pc + pw / 2
    tx1, tx2 = tc - tw / 2, tc + tw / 2
    iw = (torch.min(px2, tx2) - torch.max(px1, tx1)).clip(min=0)
    ih = (torch.min(ph, th) - torch.zeros_like(ph)).clip(min=0)
    inter = iw * ih
    union = pw * ph + tw * th - inter
    iou = inter / (union + e)
    cw = torch.max(px2, tx2) - torch.min(px1, tx1)
    ch = torch.max(ph, th)
    c2 = cw.pow(2) + ch.pow(2) + e
    v = (4 / (torch.pi ** 2)) * (torch.arctan(tw / th) - torch.arctan(pw / ph)).pow(2)
    rho2 = (pc - tc).pow(2) + (ph / 2 - th / 2).pow(2)
    a = (v / ((1 + e) - iou) + v).detach()
    return (iou - (rho2 / c2 + a * v)).squeeze(-1).clip(min=0)


def detection_loss(preds: Sequence[Tensor], targets: Tensor, anchors_dict: Dict[str, List[float]], num_classes: int,
                   anchor_t: float = 5, edge_t: float = 0.5, sample_duration: float = 60, box_w: float = 0.1,
                   conf_w: float = 1.0, class_w: float = 0.3, label_smoothing: float = 0.08,
                   ignore_index: int = -100, multi_label: bool = True, class_weights: Optional[Tensor] = None,
                   alpha: Optional[float] = None, gamma: Optional[float] = None) -> Tuple[Tensor, Dict[str, float]]:
    """AudioDetectionLoss.forward.  ref: modules/_loss.py:83-190; defaults = the reference train_config (multi_label BCE class
    loss, BCEWithLogits objectness).  multi_label=False: CrossEntropyLoss(weight=class_weights) over the class-valid matches
    (:79-81,157-158); alpha and gamma: FocalLoss(with_logits=True) objectness = mean(alpha (1 - exp(-bce))^gamma bce) (:9-37,
    74-75).  Metrics: the device-computable subset."""
    lbox = lconf = lcls = 0.0
    ws = (4.0, 2.0, 1.0)
    met = {"mean_ciou": 0.0, "conf_loss": 0.0, "class_loss": 0.0, "avg_pos_conf": 0.0, "avg_neg_conf": 0.0}
    for p, name, w in zip(preds, ("sm", "md", "lg"), ws):
        (bi, gi, ai), cl, cw = build_target_by_scale(targets, p.shape[1], anchors_dict[name], anchor_t,
                                                     sample_duration, edge_t)
        mp = p[bi, gi, ai]
        ciou = compute_ciou(mp[:, -2:], cw)
        box = (1 - ciou).mean()
        t_conf = torch.zeros(p.shape[:-1], dtype=p.dtype)
        # duplicate (b,g,a) keys (Q12): torch's non-accumulating index_put_ is "last writer wins" while it runs serially
        # (small M, e.g. the golden fixture) and order-undefined once it parallelises (M > ~4096 on CPU, always on CUDA).
        # The oracle pins the serial rule explicitly: the LAST match of a cell owns t_conf.
        keys = ((bi * p.shape[1] + gi) * p.shape[2] + ai).numpy()
        _, first_rev = np.unique(keys[::-1], return_index=True)
        last = torch.from_numpy(np.sort(len(keys) - 1 - first_rev).astype(np.int64))
        t_conf[bi[last], gi[last], ai[last]] = ciou.detach()[last]
        if alpha and gamma:
            bce = F.binary_cross_entropy_with_logits(p[..., 0], t_conf, reduction="none")
            conf = ((alpha * (1 - torch.exp(-bce)) ** gamma) * bce).mean()
        else:
            conf = F.binary_cross_entropy_with_logits(p[..., 0], t_conf)
        m = cl != ignore_index
        pcls = mp[:, 1:1 + num_classes][m]
        if multi_label:
            cn = 0.5 * label_smoothing
            tcls = torch.full_like(pcls, cn)
            tcls[range(int(m.sum())), cl[m]] = 1.0 - cn
            cls = F.binary_cross_entropy_with_logits(pcls, tcls)
        else:
            cls = F.cross_entropy(pcls, cl[m], weight=class_weights, ignore_index=ignore_index)
        hn = lambda v: v if bool(v == v) else torch.tensor(0.0)
        lbox = lbox + hn(box)
        lconf = lconf + w * hn(conf)
        lcls = lcls + hn(cls)
        met["mean_ciou"] += float(ciou.mean()) / 3
        met["conf_loss"] += float(conf) / 3
        met["class_loss"] += float(cls) / 3
        met["avg_pos_conf"] += float(p[..., 0][bi, gi, ai].sigmoid().mean()) / 3
        met["avg_neg_conf"] += float(p[..., 0][t_conf == 0].sigmoid().mean()) / 3
    loss = box_w * lbox + conf_w * lconf + class_w * lcls
    met["aggregate_loss"] = float(loss)
    return loss, met


def ema_momentum(n: int, momentum: float = 0.002, N: int = 2000) -> float:
    """ref: smoothener/_ema.py:15 (warm-up momentum, starts at 1.0)."""
    return 1 - ((1 - momentum) * (1 - math.exp(-n / N)))


def ema_update(ema: Sequence[Tensor], params: Sequence[Tensor], n: int, momentum: float = 0.002, N: int = 2000):
    """ref: smoothener/_ema.py:20-26 (parameters only)."""
    m = ema_momentum(n, momentum, N)
    for e, p in zip(ema, params):
        e.mul_(1 - m).add_(p, alpha=m)


def adam_step(params, grads, exp_avg, exp_avg_sq, step: int, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
              weight_decay=0.002):
    """torch.optim.Adam single-tensor math (L2 weight decay, not AdamW).

    ref: train.py:83-90 + config/config.yaml:75-80 -> torch/optim/adam.py (_single_tensor_adam)."""
    b1, b2 = betas
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g + weight_decay * p
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))


# =====================================================================================================================
# Anchor clustering.  ref: compute_anchors.py:63-87 -> sklearn.cluster.KMeans (sklearn 1.9.0, the image's
# version; not vendored under /root/reference), algorithm "lloyd" on the 1-D segment durations, 9 clusters, sorted centres cut
# into sm / md / lg triples.  Restated from sklearn/cluster/_kmeans.py (_kmeans_plusplus, _kmeans_single_lloyd, _tolerance,
# KMeans.fit) and _k_means_lloyd.pyx / _k_means_common.pyx (E-step argmin of |c|^2 - 2 x.c, empty-cluster relocation, averaging).
# Pinned by tests/golden/anchors.npz (the live sklearn run the way compute_anchors.py runs it: global numpy RNG seeded with 42).
def _sq_dists_1d(xc: np.ndarray, x: np.ndarray, xsq: np.ndarray) -> np.ndarray:
    """sklearn.metrics.pairwise._euclidean_distances(X[cand], X, squared=True) for one feature: -2 x.y + |x|^2 + |y|^2, clipped."""
    d = -2.0 * (xc[:, None] * x[None, :])
    d += (xc * xc)[:, None]
    d += xsq[None, :]
    np.maximum(d, 0, out=d)
    return d


def kmeans_plusplus_init(x: np.ndarray, k: int, rng: np.random.RandomState) -> np.ndarray:
    """sklearn _kmeans_plusplus (unit sample weights) on centred 1-D data x [n]; returns the k initial centres."""
    n = x.shape[0]
    xsq = x * x
    w = np.ones(n)
    n_local_trials = 2 + int(np.log(k))
    centers = np.empty(k)
    cid = rng.choice(n, p=w / w.sum())
    centers[0] = x[cid]
    closest = _sq_dists_1d(centers[:1], x, xsq)
    pot = closest @ w
    for c in range(1, k):
        rand_vals = rng.uniform(size=n_local_trials) * pot
        cand = np.searchsorted(np.cumsum(w * closest), rand_vals)
        np.clip(cand, None, closest.size - 1, out=cand)
        d = _sq_dists_1d(x[cand], x, xsq)
        np.minimum(closest, d, out=d)
        cpot = d @ w.reshape(-1, 1)
        best = int(np.argmin(cpot))
        pot = cpot[best]
        closest = d[best]
        centers[c] = x[cand[best]]
    return centers


def kmeans_lloyd_1d(x: np.ndarray, centers_init: np.ndarray, max_iter: int, tol_abs: float):
    """sklearn _kmeans_single_lloyd on centred 1-D data: returns (labels, inertia, centres, n_iter)."""
    k = centers_init.shape[0]
    centers = centers_init.astype(np.float64).copy()
    labels_old = np.full(x.shape[0], -1, np.int64)
    strict = False
    it = 0

    def e_step(c):
        return np.argmin((c * c)[None, :] - 2.0 * (x[:, None] * c[None, :]), axis=1)     # first minimum wins ties

    for it in range(max_iter):
        labels = e_step(centers)
        sums = np.bincount(labels, weights=x, minlength=k)
        cnt = np.bincount(labels, minlength=k).astype(np.float64)
        empty = np.where(cnt == 0)[0]
        if empty.size:                               # _relocate_empty_clusters_dense: farthest points become the new centres
            dist = (x - centers[labels]) ** 2
            far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
            for e, f in zip(empty, far):
                sums[labels[f]] -= x[f]
                cnt[labels[f]] -= 1
                sums[e] = x[f]
                cnt[e] = 1
        new = sums / cnt
        shift_tot = float(((new - centers) ** 2).sum())
        centers = new
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift_tot <= tol_abs:
            break
        labels_old = labels
    if not strict:
        labels = e_step(centers)
    inertia = float(((x - centers[labels]) ** 2).sum())
    return labels, inertia, centers, it + 1


def compute_anchors(durations: Sequence[float], n_clusters: int = 9, init: str = "k-means++", n_init="auto", max_iter: int = 500,
                    tol: float = 1e-10, rng: Optional[np.random.RandomState] = None):
    """compute_anchors.py:72-86: KMeans(9, init, n_init, tol, max_iter).fit(durations) -> sorted centres -> (sm, md, lg)."""
    rng = rng if rng is not None else np.random.mtrand._rand
    X = np.asarray(durations, np.float64).reshape(-1)
    tol_abs = float(np.var(X)) * tol
    mean = X.mean()
    x = X - mean
    if n_init == "auto":
        n_init = 1 if init == "k-means++" else 10
    best = None
    for _ in range(int(n_init)):
        if init == "k-means++":
            c0 = kmeans_plusplus_init(x, n_clusters, rng)
        else:
            c0 = x[rng.choice(x.shape[0], size=n_clusters, replace=False, p=np.ones(x.shape[0]) / x.shape[0])]
        labels, inertia, centers, n_iter = kmeans_lloyd_1d(x, c0, max_iter, tol_abs)
        if best is None or inertia < best[1]:
            best = (labels, inertia, centers, n_iter)
    a = np.sort(best[2] + mean)
    return a[:3], a[3:6], a[6:], best
