"""Parity of the sm_100a kernels against the CPU oracle and the reference-made golden fixtures.
All calls go through the C ABI (ctypes) - via the host package or directly.  Run with -m gpu on a B200."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth
import yad_b200
from oracle import ref_port as O
from yad_b200 import _lib
from yad_b200._lib import ACT_LRELU, ACT_NONE, ACT_RELU, BF16, F32, ConvDesc

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.fixture(scope="module")
def models(cuda_dev, ref_state_dict):
    out = {}
    for form in ("train", "deploy"):
        for dt in ("f32", "bf16"):
            m = yad_b200.AudioDetectionNetwork(2, compute_dtype=dt)
            m.load_state_dict(ref_state_dict)
            if form == "deploy":
                m.inference()
            out[(form, dt)] = m.eval().to(cuda_dev)
    return out


# ------------------------------------------------------------------ frontend (S1)
def _check_frontend(taps, fe, atol_ch0=2e-4):
    mel, rmel = taps["mel"].cpu().numpy(), fe["mel"].numpy()
    np.testing.assert_allclose(mel, rmel, rtol=2e-4, atol=1e-6)            # fp32 FFT vs MKL rfft
    np.testing.assert_allclose(taps["meldb"].cpu().numpy(), fe["meldb"].numpy(), atol=2e-3)
    np.testing.assert_allclose(taps["mfcc"].cpu().numpy(), fe["mfcc"].numpy(), atol=5e-3)   # pre-dB MFCC
    xs, rxs = taps["x_spectral"].cpu().numpy(), fe["x_spectral"].numpy()
    np.testing.assert_allclose(xs[:, 0], rxs[:, 0], atol=atol_ch0)          # standardised dB-mel
    # channel 1 = dB of signed MFCCs: log of values that cross zero (fp64-vs-fp32 of the SAME algorithm already
    # moves 17 % of this plane by > 1e-3), so: tight where the MFCC is away from zero, loose in the mean elsewhere.
    d = np.abs(xs[:, 1] - rxs[:, 1])
    away = np.abs(fe["mfcc"].numpy()[:, 0]) > 0.05
    assert d[away].max() < 0.05, d[away].max()
    assert d.mean() < 2e-2, d.mean()


def test_frontend_short_clips_vs_oracle_and_golden(models, gold, ref_state_dict, cuda_dev):
    x = synth.synth_clips(3, 22050 * 6, seed=1000, silence_tail_every=3)
    taps = {}
    models[("train", "f32")](x.to(cuda_dev), combine_scales=True, taps=taps)
    _check_frontend(taps, O.frontend(x, ref_state_dict))
    g = gold("short_clips")
    np.testing.assert_allclose(taps["mel"].cpu().numpy(), g["mel"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(taps["x_spectral"].cpu().numpy()[:, 0], g["x_spectral"][:, 0], atol=2e-4)


def test_frontend_full_clip_and_ragged_lengths(models, gold, ref_state_dict, cuda_dev):
    m = models[("train", "f32")]
    x = synth.synth_clips(1, 1323000, seed=2000, silence_tail_every=0)
    taps = {}
    m(x.to(cuda_dev), combine_scales=True, taps=taps)
    assert taps["x_spectral"].shape == (1, 2, 32, 960)
    np.testing.assert_allclose(taps["x_spectral"].cpu().numpy()[:, 0], gold("full_clip")["x_spectral"][:, 0], atol=2e-4)
    # ragged: a length that is neither a whole number of hops nor of 8-frame groups (T = 61)
    xr = synth.synth_clips(2, 84321, seed=3000, silence_tail_every=2)
    taps = {}
    m(xr.to(cuda_dev), combine_scales=True, taps=taps)
    _check_frontend(taps, O.frontend(xr, ref_state_dict))
    # an ODD frame count above the cluster threshold of stage B (T = 125: two CTAs per clip with 63 and 62 frames; the maxima and
    # moments cross the cluster through distributed shared memory - modules/_architecture.py:98-105, 182-189)
    xo = synth.synth_clips(3, 173000, seed=3100, silence_tail_every=3)
    taps = {}
    m(xo.to(cuda_dev), combine_scales=True, taps=taps)
    assert taps["x_spectral"].shape == (3, 2, 32, 125)
    _check_frontend(taps, O.frontend(xo, ref_state_dict))


# ---- the benchmarked code path: persistent CTAs running MANY groups each (the bench has ~415 groups per CTA; every test above
#      has at most one).  9 full-length clips = 1080 groups -> 8 per CTA on 148 SMs: double-buffered staging, the BAR_EMPTY
#      handshake (gi >= 2), mbarrier phase flips and clip-boundary crossings all run.  modules/_architecture.py:84-108.
@pytest.fixture(scope="module")
def long_batch():
    return synth.synth_clips(9, 1323000, seed=5000, silence_tail_every=3)


@pytest.mark.parametrize("kind", ["f32", "int16", "taper"])
def test_frontend_multigroup_persistent_path(models, variant_state_dict, ref_state_dict, long_batch, kind, cuda_dev):
    x = long_batch
    cfg = O.DEFAULT_CONFIG
    if kind == "taper":
        sd, cfg = variant_state_dict("taper")
        m = yad_b200.AudioDetectionNetwork(2, config=cfg, compute_dtype="f32")
        m.load_state_dict(sd)
        m = m.eval().to(cuda_dev)
    else:
        m = models[("train", "f32")]
    if kind == "int16":
        xi = (x.clamp(-1, 1) * 32767).round().to(torch.int16)
        x, xin = xi.float() / 32768.0, xi
    else:
        xin = x
    taps = {}
    m(xin.to(cuda_dev), combine_scales=True, taps=taps)
    assert taps["mel"].shape == (9, 1, 32, 960)
    _check_frontend(taps, O.frontend(x, ref_state_dict, cfg))
    # size-independent property of the persistent schedule: a clip's planes do not depend on where in the flat group sequence
    # (which CTA, which staging buffer, before / after a clip boundary) its groups ran - bitwise
    for idx in ([0], [4, 2], [8, 7, 6]):
        t2 = {}
        m(xin[idx].to(cuda_dev), combine_scales=True, taps=t2)
        assert torch.equal(t2["mel"], taps["mel"][idx]), (kind, idx)
        assert torch.equal(t2["x_spectral"], taps["x_spectral"][idx]), (kind, idx)


# ------------------------------------------------------------------ convolutions
CONV_CASES = [
    # B, H, W, Cin, Cout, k, stride, pad, act, residual
    (2, 8, 48, 64, 64, 3, (1, 1), 1, ACT_RELU, True),       # layer1
    (2, 16, 96, 64, 64, 7, (2, 2), 3, ACT_RELU, False),     # stem conv2
    (3, 8, 48, 64, 128, 3, (2, 2), 1, ACT_RELU, False),     # layer2.0.conv1
    (3, 8, 48, 64, 128, 1, (2, 2), 0, ACT_NONE, False),     # downsample 1x1 s2
    (2, 2, 12, 256, 256, 3, (1, 1), 1, ACT_RELU, True),     # layer3
    (5, 1, 6, 512, 512, 3, (1, 1), 1, ACT_RELU, True),      # layer4 (H = 1: top/bottom taps are padding)
    (4, 1, 24, 128, 15, 3, (1, 1), 1, ACT_LRELU, False),    # RepVGG 128 -> 15 head
    (4, 1, 24, 15, 128, 3, (1, 2), 1, ACT_LRELU, False),    # conv2_downsample (stride (1,2), Cin = 15)
    (4, 1, 12, 512, 64, 1, (1, 1), 0, ACT_LRELU, False),    # CSPSPPF 1x1
    (1, 8, 240, 64, 64, 3, (1, 1), 1, ACT_RELU, True),      # full-width layer1 row tiles
]


def _conv_ref(x, w, b, stride, pad, act, res):
    y = F.conv2d(x, w, b, stride=stride, padding=pad)
    if res is not None:
        y = y + res
    if act == ACT_RELU:
        y = F.relu(y)
    elif act == ACT_LRELU:
        y = F.leaky_relu(y, 0.2)
    return y


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("path", ["simt_f32", "tc_bf16"])
def test_conv_kernels(case, path, cuda_dev):
    B, H, W, Cin, Cout, k, stride, pad, act, use_res = case
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(hash(case) % 10000)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    Ho, Wo = (H + 2 * pad - k) // stride[0] + 1, (W + 2 * pad - k) // stride[1] + 1
    res = torch.randn(B, Cout, Ho, Wo, generator=g) if use_res else None
    bf = path == "tc_bf16"
    if bf:   # the oracle sees the same bf16-rounded operands; accumulation differs (fp32 both)
        x, w = x.bfloat16().float(), w.bfloat16().float()
        res = res.bfloat16().float() if use_res else None
    ref = _conv_ref(x, w, b, stride, pad, act, res)
    td = torch.bfloat16 if bf else torch.float32
    cin_p = (Cin + 63) // 64 * 64 if bf else Cin
    ld_out = (Cout + 63) // 64 * 64 + 64 if bf else Cout + 3         # exercise pitch != channels and a slice offset
    co_off = 64 if bf else 2
    xin = torch.zeros(B, H, W, cin_p, dtype=td, device=cuda_dev)
    xin[..., :Cin] = x.permute(0, 2, 3, 1).to(td)
    out = torch.full((B, Ho, Wo, ld_out), 7.0, dtype=td, device=cuda_dev)
    rin = res.permute(0, 2, 3, 1).contiguous().to(td).to(cuda_dev) if use_res else None
    if bf and use_res:
        rp = torch.zeros(B, Ho, Wo, (Cout + 7) // 8 * 8, dtype=td, device=cuda_dev); rp[..., :Cout] = rin; rin = rp
    d = ConvDesc(B=B, H=H, W=W, Cin=cin_p, ld_in=cin_p, Cout=Cout, ld_out=ld_out, co_off=co_off, kh=k, kw=k, sh=stride[0],
                 sw=stride[1], ph=pad, pw=pad, act=act, ld_res=rin.shape[3] if use_res else 0)
    bias = b.to(cuda_dev)
    if bf:
        cout_p = (Cout + 15) // 16 * 16
        wt = torch.zeros(cout_p, k, k, cin_p)
        wt[:Cout, :, :, :Cin] = w.permute(0, 2, 3, 1)
        wt = wt.reshape(cout_p, -1).to(td).to(cuda_dev)
        bp = torch.zeros(cout_p, device=cuda_dev); bp[:Cout] = bias
        out2 = torch.zeros(B, Ho, Wo, (Cout + 3) // 4 * 4, device=cuda_dev)
        rc = lib.yad_conv_tc(C.byref(d), xin.data_ptr(), wt.data_ptr(), cout_p, bp.data_ptr(), _lib.ptr(rin), out.data_ptr(), BF16,
                             out2.data_ptr(), out2.shape[3], _stream())
    else:
        wt = w.permute(2, 3, 1, 0).contiguous().to(cuda_dev)
        rc = lib.yad_conv_simt(C.byref(d), F32, xin.data_ptr(), wt.data_ptr(), Cout, bias.data_ptr(), _lib.ptr(rin), out.data_ptr(),
                               _stream())
    _lib.check(rc, path)
    torch.cuda.synchronize()
    got = out[..., co_off:co_off + Cout].float().permute(0, 3, 1, 2).cpu()
    if bf:
        # fp32 copy: only accumulation order differs from the oracle -> tight; bf16 output: + one rounding (2^-8 rel)
        got2 = out2[..., :Cout].permute(0, 3, 1, 2).cpu()
        np.testing.assert_allclose(got2.numpy(), ref.numpy(), atol=2e-3, rtol=2e-3)
        np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-2, rtol=1e-2)
    else:
        np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-4, rtol=1e-4)
    # channels outside the written slice are untouched
    assert torch.all(out[..., :co_off].float() == 7.0) and torch.all(out[..., co_off + Cout:].float() == 7.0)


def test_stem_conv(cuda_dev):
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 2, 32, 96, generator=g)
    w = torch.randn(64, 2, 7, 7, generator=g) * 0.1
    ref = F.conv2d(x, w, None, stride=2, padding=3)
    wt = w.permute(2, 3, 1, 0).contiguous().to(cuda_dev)
    for dt, td, tol in ((F32, torch.float32, 1e-4), (BF16, torch.bfloat16, 2e-2)):
        out = torch.empty(2, 16, 48, 64, dtype=td, device=cuda_dev)
        _lib.check(lib.yad_conv_stem(x.to(cuda_dev).data_ptr(), 2, 32, 96, wt.data_ptr(), out.data_ptr(), dt, _stream()), "stem")
        np.testing.assert_allclose(out.float().permute(0, 3, 1, 2).cpu().numpy(), ref.numpy(), atol=tol, rtol=tol)


@pytest.mark.parametrize("shape", [(2, 32, 96), (1, 32, 960), (3, 32, 61), (2, 7, 300)])
def test_stem_conv_tensor_core(shape, cuda_dev):
    """tcgen05 stem vs F.conv2d on the same bf16-rounded operands (fp32 accumulate both; one bf16 output rounding)."""
    lib = _lib.init(0)
    B, H, W = shape
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, 2, H, W, generator=g).bfloat16().float()
    w = (torch.randn(64, 2, 7, 7, generator=g) * 0.1).bfloat16().float()
    ref = F.conv2d(x, w, None, stride=2, padding=3)
    wk = torch.zeros(64, 7, 16)
    wk[:, :, :14] = w.permute(0, 2, 3, 1).reshape(64, 7, 14)
    wp = wk.reshape(8, 8, 14, 8).permute(0, 2, 1, 3).contiguous().to(torch.bfloat16).to(cuda_dev)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.full((B, Ho, Wo, 64), 7.0, dtype=torch.bfloat16, device=cuda_dev)
    _lib.check(lib.yad_conv_stem_tc(x.to(cuda_dev).data_ptr(), B, H, W, wp.data_ptr(), out.data_ptr(), 0, 0, _stream()), "stem_tc")
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.float().permute(0, 3, 1, 2).cpu().numpy(), ref.numpy(), atol=2e-2, rtol=1e-2)


FLAT_CASES = [
    # B, H, W, Cin, Cout, k, act, residual
    (2, 8, 48, 64, 64, 3, ACT_RELU, True),        # layer1
    (3, 4, 24, 128, 128, 3, ACT_RELU, True),      # layer2
    (2, 2, 12, 256, 256, 3, ACT_RELU, True),      # layer3
    (5, 1, 6, 512, 512, 3, ACT_RELU, True),       # layer4 (H = 1: rows -1 / +1 of the filter only see padding)
    (1, 8, 240, 64, 64, 3, ACT_RELU, False),      # full-width layer1
    (4, 1, 24, 128, 128, 3, ACT_LRELU, False),    # neck RepVGG (deploy form)
    (2, 4, 24, 64, 128, 1, ACT_NONE, False),      # 1x1
    (37, 8, 30, 64, 64, 3, ACT_RELU, True),       # many super-tiles per CTA, ragged last tile
]


def _to_flat(x, Hp, Wp, ld, dev):
    """NCHW fp32 -> flat halo-padded bf16 [B, Wp, Hp, ld] (pixel (b,h,w) at (b*Wp + w)*Hp + h)."""
    B, Cc, H, W = x.shape
    buf = torch.zeros(B, Wp, Hp, ld, dtype=torch.bfloat16)
    buf[:, :W, :H, :Cc] = x.permute(0, 3, 2, 1).to(torch.bfloat16)
    return buf.to(dev)


@pytest.mark.parametrize("case", FLAT_CASES)
def test_conv_flat(case, cuda_dev):
    """Patch-resident tcgen05 conv on the flat layout vs F.conv2d on the same bf16-rounded operands."""
    B, H, W, Cin, Cout, k, act, use_res = case
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(hash(case) % 10000)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(Cout, generator=g) * 0.1
    res = torch.randn(B, Cout, H, W, generator=g).bfloat16().float() if use_res else None
    ref = _conv_ref(x, w, b, 1, k // 2, act, res)
    Hp, Wp = (H + k // 2 if H > 1 else 1), W + k // 2
    ld_out, co_off = Cout + 64, 64
    xin = _to_flat(x, Hp, Wp, Cin, cuda_dev)
    rin = _to_flat(res, Hp, Wp, Cout, cuda_dev) if use_res else None
    out = torch.full((B, Wp, Hp, ld_out), 7.0, dtype=torch.bfloat16, device=cuda_dev)
    cout_p = (Cout + 63) // 64 * 64
    wt = torch.zeros(cout_p, k, k, Cin)
    wt[:Cout] = w.permute(0, 2, 3, 1)
    wt = wt.reshape(cout_p, -1).to(torch.bfloat16).to(cuda_dev)
    bp = torch.zeros(cout_p, device=cuda_dev); bp[:Cout] = b.to(cuda_dev)
    d = _lib.FlatDesc(B=B, H=H, W=W, Hp=Hp, Wp=Wp, Cin=Cin, ld_in=Cin, Cout=Cout, ld_out=ld_out, co_off=co_off, kh=k, kw=k,
                      ph=k // 2, pw=k // 2, act=act, ld_res=Cout if use_res else 0)
    flags = 0
    _lib.check(lib.yad_conv_flat(C.byref(d), xin.data_ptr(), wt.data_ptr(), cout_p, bp.data_ptr(), _lib.ptr(rin), out.data_ptr(),
                                 flags, _stream()), "conv_flat")
    torch.cuda.synchronize()
    got = out[:, :W, :H, co_off:co_off + Cout].float().permute(0, 3, 2, 1).cpu()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-2, rtol=1e-2)
    # channels outside the slice are never written; halo cells of the slice are either left alone (register-sourced epilogue)
    # or written as the zeros they have to be (whole tiles stored by the TMA unit)
    halo = torch.cat([out[:, W:, :, co_off:co_off + Cout].reshape(-1), out[:, :, H:, co_off:co_off + Cout].reshape(-1)]).float()
    assert torch.all((halo == 7.0) | (halo == 0.0))
    assert torch.all(out[..., :co_off].float() == 7.0) and torch.all(out[..., co_off + Cout:].float() == 7.0)


# ------------------------------------------------------------------ whole network (S2/S3)
@pytest.mark.parametrize("form", ["train", "deploy"])
def test_network_f32_vs_golden(models, gold, form, cuda_dev):
    g = gold("short_clips")
    x = synth.synth_clips(3, 22050 * 6, seed=1000, silence_tail_every=3).to(cuda_dev)
    taps = {}
    out = models[(form, "f32")](x, combine_scales=True, taps=taps)
    assert out.shape == (3, 63, 5)
    pre = "head" if form == "train" else "dhead"
    for h, n in zip(taps["heads"], ("sm", "md", "lg")):
        np.testing.assert_allclose(h.cpu().numpy(), g[f"{pre}_{n}"], atol=5e-3)     # fp32 logits / offsets
    np.testing.assert_allclose(out.cpu().numpy(), g[f"preds_{form}"], atol=5e-3)
    if form == "train":
        np.testing.assert_allclose(taps["fmaps"][0].cpu().numpy(), g["fmap1"], atol=2e-3)
        np.testing.assert_allclose(taps["fmaps"][3].cpu().numpy(), g["fmap4"], atol=2e-3)
    sm, md, lg = models[(form, "f32")](x)                  # tuple form, shapes of the reference
    assert sm.shape == (3, 12, 3, 5) and md.shape == (3, 6, 3, 5) and lg.shape == (3, 3, 3, 5)
    np.testing.assert_array_equal(torch.cat([sm.reshape(3, -1, 5), md.reshape(3, -1, 5), lg.reshape(3, -1, 5)], 1).cpu().numpy(),
                                  out.cpu().numpy())


@pytest.mark.parametrize("form", ["train", "deploy"])
def test_network_bf16_vs_golden(models, gold, form, cuda_dev):
    """bf16 tensor-core path: stated tolerance on logits 0.15 abs / mean 0.02 (SURVEY Q14 calibration: CPU bf16
    autocast vs fp32 gives max 0.048 / mean 0.006 on the deploy form), centres 0.1 s, widths 1.5 s max."""
    g = gold("short_clips")
    x = synth.synth_clips(3, 22050 * 6, seed=1000, silence_tail_every=3).to(cuda_dev)
    out = models[(form, "bf16")](x, combine_scales=True).cpu().numpy()
    ref = g[f"preds_{form}"]
    d = np.abs(out - ref)
    assert d[..., :3].max() < 0.15 and d[..., :3].mean() < 0.02, (d[..., :3].max(), d[..., :3].mean())
    assert d[..., 3].max() < 0.1, d[..., 3].max()
    assert d[..., 4].max() < 1.5 and d[..., 4].mean() < 0.15, (d[..., 4].max(), d[..., 4].mean())


def test_full_clip_config1_end_to_end(models, gold, cuda_dev):
    g = gold("full_clip")
    x = synth.synth_clips(1, 1323000, seed=2000, silence_tail_every=0).to(cuda_dev)
    out = models[("train", "f32")](x, combine_scales=True)
    np.testing.assert_allclose(out.cpu().numpy(), g["preds_train"], atol=5e-3)
    outd = models[("deploy", "f32")](x, combine_scales=True)
    np.testing.assert_allclose(outd.cpu().numpy(), g["preds_deploy"], atol=5e-3)


def _bf16_close(out, ref, what=""):
    """The stated bf16 tolerance of the tcgen05 path (see test_network_bf16_vs_golden)."""
    d = np.abs(out - ref)
    assert d[..., :3].max() < 0.15 and d[..., :3].mean() < 0.02, (what, d[..., :3].max(), d[..., :3].mean())
    assert d[..., 3].max() < 0.1, (what, d[..., 3].max())
    assert d[..., 4].max() < 1.5 and d[..., 4].mean() < 0.15, (what, d[..., 4].max(), d[..., 4].mean())


def test_bf16_deploy_full_clip_vs_live_reference(models, gold, cuda_dev):
    """BASELINE configs[2] shape per clip (T = 960, W = 240 -> multi-super-tile flat convs) on the bf16 tcgen05 path against the
    LIVE reference's deploy-form prediction of the 60 s fixture (tests/golden/full_clip.npz)."""
    g = gold("full_clip")
    x = synth.synth_clips(1, 1323000, seed=2000, silence_tail_every=0).to(cuda_dev)
    out = models[("deploy", "bf16")](x, combine_scales=True).cpu().numpy()
    assert out.shape == g["preds_deploy"].shape == (1, 630, 5)
    _bf16_close(out, g["preds_deploy"], "deploy bf16 T=960")
    outt = models[("train", "bf16")](x, combine_scales=True).cpu().numpy()
    _bf16_close(outt, g["preds_train"], "train-form bf16 T=960")


def test_bf16_deploy_batch160_vs_oracle(models, ref_state_dict, cuda_dev):
    """More clips than SMs (B = 160 > 148: every per-clip kernel and every persistent tile loop wraps) at full length: 4 sampled
    clips against the oracle at the bf16 tolerance, and every replica of a clip bitwise equal to its first occurrence (batch
    position must not change a result), and equal to the same clip run in a batch of 4."""
    m = models[("deploy", "bf16")]
    base = synth.synth_clips(16, 1323000, seed=6000, silence_tail_every=8)
    idx = torch.arange(160) % 16
    x = base[idx].to(cuda_dev)
    out = m(x, combine_scales=True)
    assert out.shape == (160, 630, 5)
    for i in range(16, 160):
        assert torch.equal(out[i], out[i % 16]), i
    pick = [0, 37, 148, 159]
    small = m(x[pick].contiguous(), combine_scales=True)
    assert torch.equal(small, out[pick])
    sd = O.fold_repvgg({k: v.cpu() for k, v in ref_state_dict.items()})
    ref = O.forward(base[[p % 16 for p in pick]], sd, 2)
    _bf16_close(out[pick].cpu().numpy(), ref.numpy(), "B=160")
    # the post-processing of the whole batch: clip ids of replicas repeat with period 16
    seg, bidx = yad_b200.process_model_outputs(out, 0.1, 0.2)
    cnt = torch.bincount(bidx.cpu(), minlength=160)
    assert torch.equal(cnt[:16].repeat(10), cnt)


def test_decode_vs_oracle(cuda_dev, ref_state_dict):
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(4)
    heads = [torch.randn(2, G, 15, generator=g) * 3 for G in (120, 60, 30)]
    ref = O.decode(heads, ref_state_dict, 960000, 960, 2)
    hd = [h.to(cuda_dev).contiguous() for h in heads]
    preds = torch.empty(2, 630, 5, device=cuda_dev)
    hp = (C.c_void_p * 3)(*[h.data_ptr() for h in hd])
    anc = torch.cat([ref_state_dict[f"{n}_anchors"] * 60 for n in ("sm", "md", "lg")]).tolist()
    rc = lib.yad_decode(hp, (C.c_int32 * 3)(120, 60, 30), (C.c_int32 * 3)(15, 15, 15), (C.c_int32 * 3)(8, 16, 32), 3, F32,
                        (C.c_float * 9)(*anc), 3, 2, 16.0, 60.0, 2, preds.data_ptr(), _stream())
    _lib.check(rc, "decode")
    got = preds.cpu().numpy()
    np.testing.assert_array_equal(got[..., :3], ref.numpy()[..., :3])         # logits pass through untouched
    np.testing.assert_allclose(got[..., 3:], ref.numpy()[..., 3:], rtol=3e-6, atol=1e-5)   # few ulp (sigmoid)


# ------------------------------------------------------------------ NMS (S4) - bit exact
def _canon(seg, bidx):
    """The reference orders a clip's segments with an argsort that is not declared stable (inference.py:95), so rows
    whose sort key (centre) ties - the planted zero-width twins - have no defined order: sort such runs by confidence."""
    seg, bidx = np.array(seg), np.asarray(bidx)
    i = 0
    while i < len(seg):
        j = i + 1
        while j < len(seg) and bidx[j] == bidx[i] and seg[j, 3] == seg[i, 3] and seg[j, 4] == seg[i, 4]:
            j += 1
        if j - i > 1:
            seg[i:j] = seg[i:j][np.argsort(seg[i:j, 0], kind="stable")]
        i = j
    return seg


def _check_nms(o, iou, cthr, gold_seg=None, gold_bidx=None):
    r = yad_b200.nms_raw(o.cuda(), iou, cthr, want_taps=True)
    conf, boxes = r["conf"].cpu(), r["boxes"].cpu()
    rcoords, rconf = O.boxes_and_confidence(o)
    np.testing.assert_array_equal(boxes.numpy(), rcoords[..., [0, 2]].numpy())          # fp32 box arithmetic: exact
    np.testing.assert_allclose(conf.numpy(), rconf.numpy(), rtol=2e-6, atol=1e-9)        # exp() ulp differences only
    B, P = conf.shape
    keep, nk = r["keep"].cpu().numpy(), r["n_keep"].cpu().numpy()
    for b in range(B):
        # NMS core exact given identical (boxes, conf): oracle greedy NMS on the kernel's own confidences
        c4 = torch.stack([boxes[b, :, 0], torch.zeros(P), boxes[b, :, 1], torch.full((P,), 10.0)], 1)
        want = O.nms_greedy(c4.numpy(), conf[b].numpy(), iou)
        np.testing.assert_array_equal(keep[b, :nk[b]], want)
        assert np.all(keep[b, nk[b]:] == -1)
    return r


@pytest.mark.parametrize("name,B", [("b1", 1), ("b3", 3)])
@pytest.mark.parametrize("iou,cthr", [(0.1, 0.2), (0.1, 0.65), (0.05, 0.5)])
def test_nms_keep_sets_and_segments(gold, name, B, iou, cthr, cuda_dev):
    g = gold("nms")
    o = synth.synth_heads(B, 630, 2, seed=7 + B)
    r = _check_nms(o, iou, cthr)
    ref = g[f"keep_{name}_{iou}_{cthr}"]                       # torchvision's own keep list
    keep, nk = r["keep"].cpu().numpy(), r["n_keep"].cpu().numpy()
    for b in range(B):
        np.testing.assert_array_equal(keep[b, :nk[b]] + b * 630, ref[(ref >= b * 630) & (ref < (b + 1) * 630)])
    seg, bidx = yad_b200.process_model_outputs(o.cuda(), iou, cthr)
    np.testing.assert_array_equal(bidx.cpu().numpy(), g[f"bidx_{name}_{iou}_{cthr}"])
    np.testing.assert_allclose(_canon(seg.cpu().numpy(), bidx.cpu().numpy()),
                               _canon(g[f"seg_{name}_{iou}_{cthr}"], g[f"bidx_{name}_{iou}_{cthr}"]), rtol=2e-6, atol=1e-7)


def test_nms_edge_cases(gold, cuda_dev):
    out = torch.from_numpy(gold("short_clips")["preds_train"])
    _check_nms(out, 0.1, 0.2)                                  # P = 63 (not a multiple of 32)
    seg, bidx = yad_b200.process_model_outputs(out.cuda(), 0.1, 0.2)
    g = gold("nms")
    np.testing.assert_array_equal(bidx.cpu().numpy(), g["bidx_model_0.1_0.2"])
    np.testing.assert_allclose(seg.cpu().numpy(), g["seg_model_0.1_0.2"], rtol=2e-6, atol=1e-7)
    with pytest.raises(ValueError):
        yad_b200.process_model_outputs(out.cuda(), 0.1, 0.999999)      # empty result raises like the reference
    seg2, b2 = yad_b200.process_model_outputs(out[0].cuda(), 0.1, 0.2)  # 2-D input accepted (inference.py:51-53)
    assert torch.all(b2 == 0) and seg2.shape[0] == int((bidx == 0).sum())
    big = synth.synth_heads(64, 630, 2, seed=99, adversarial=False)      # many clips, full-size keep-sets
    r = _check_nms(big, 0.05, 0.5)
    so, bo = O.process_model_outputs(big.clone(), 0.05, 0.5)
    sg, bg = yad_b200.process_model_outputs(big.cuda(), 0.05, 0.5)
    np.testing.assert_array_equal(bg.cpu().numpy(), bo.numpy())
    np.testing.assert_allclose(sg.cpu().numpy(), so.numpy(), rtol=2e-6, atol=1e-7)


# ------------------------------------------------------------------ training-side kernels (S5 exact, S6 tolerance)
@pytest.mark.parametrize("name,G", [("sm", 120), ("md", 60), ("lg", 30)])
def test_anchor_matching_bit_exact(gold, name, G, cuda_dev):
    g = gold("train")
    tg = torch.from_numpy(g["targets"]).cuda()
    (bi, gi, ai), cl, cw = yad_b200.build_target_by_scale(tg, G, O.DEFAULT_CONFIG["anchors"][name], 5, 60, 0.5)
    np.testing.assert_array_equal(bi.cpu().numpy(), g[f"bi_{name}"])
    np.testing.assert_array_equal(gi.cpu().numpy(), g[f"gi_{name}"])
    np.testing.assert_array_equal(ai.cpu().numpy(), g[f"ai_{name}"])
    np.testing.assert_array_equal(cl.cpu().numpy(), g[f"cl_{name}"])
    np.testing.assert_array_equal(cw.cpu().numpy(), g[f"cw_{name}"])
    big = synth.synth_targets(300, seed=5)                       # > one scan chunk (A*T > 1024)
    (b2, g2, a2), c2, w2 = yad_b200.build_target_by_scale(big.cuda(), G, O.DEFAULT_CONFIG["anchors"][name], 5, 60, 0.5)
    (b3, g3, a3), c3, w3 = O.build_target_by_scale(big, G, O.DEFAULT_CONFIG["anchors"][name], 5, 60, 0.5)
    for x, y in ((b2, b3), (g2, g3), (a2, a3), (c2, c3)):
        np.testing.assert_array_equal(x.cpu().numpy(), y.numpy())
    empty = yad_b200.build_target_by_scale(torch.zeros(0, 4).cuda(), G, O.DEFAULT_CONFIG["anchors"][name], 5, 60, 0.5)
    assert empty[1].numel() == 0


def test_fused_adam_ema(gold, cuda_dev):
    g = gold("train")
    p = torch.nn.Parameter(torch.from_numpy(g["adam_w0"]).clone().cuda())
    opt = yad_b200.FusedAdamEMA([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.002, ema_momentum=0.002, ema_N=2000,
                                use_ema=True)
    for step in range(3):
        p.grad.copy_(torch.from_numpy(g["adam_grads"][step]).cuda())
        opt.step()
    np.testing.assert_allclose(p.data.cpu().numpy(), g["adam_w3"], atol=2e-6)
    np.testing.assert_allclose(opt.ema.cpu().numpy(), g["ema3"], atol=2e-6)


# ------------------------------------------------------------------ detection loss
def _loss_module():
    lc = dict(yad_b200.default_config()["train_config"]["loss_config"])
    return yad_b200.AudioDetectionLoss(yad_b200.default_config()["anchors"], 2, **lc)


def test_detection_loss_vs_reference_golden(gold, cuda_dev):
    """Loss value and d loss / d preds of the CUDA kernels vs the values the live reference produced (fixtures)."""
    g = gold("train")
    tg = torch.from_numpy(g["targets"]).to(cuda_dev)
    with torch.enable_grad():
        preds = [torch.from_numpy(g[f"pred{i}"]).to(cuda_dev).requires_grad_(True) for i in range(3)]
        loss, met = _loss_module()(preds, tg)
        loss.backward()
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=2e-6)          # tolerance: fp32 sums in fp64 accumulators
    for i in range(3):
        np.testing.assert_allclose(preds[i].grad.cpu().numpy(), g[f"grad{i}"], atol=2e-7, rtol=2e-4)
    assert set(met) == {"aggregate_loss", "mean_ciou", "conf_loss", "avg_pos_conf", "avg_neg_conf", "class_loss", "accuracy", "f1",
                        "precision", "recall"}


LOSS_VARIANTS = {
    "ce": dict(multi_label=False),
    "ce_weighted": dict(multi_label=False, class_weights=torch.tensor([0.3, 1.7])),
    "focal": dict(multi_label=True, alpha=0.25, gamma=1.5),
    "focal_ce": dict(multi_label=False, alpha=0.4, gamma=2.0, class_weights=torch.tensor([1.2, 0.6])),
}


@pytest.mark.parametrize("name", sorted(LOSS_VARIANTS))
def test_detection_loss_variants_vs_reference_golden(gold, name, cuda_dev):
    """yad_loss_scale_ex: cross-entropy class loss (multi_label: false, with / without class weights) and the focal objectness
    loss against loss value, gradients and metrics of the LIVE reference (modules/_loss.py:9-37,74-81,157-158; fixture
    tests/golden/loss_variants.npz), plus a random case with ignore labels and duplicates against the oracle."""
    g, tr = gold("loss_variants"), gold("train")
    lc = dict(yad_b200.default_config()["train_config"]["loss_config"])
    lc.update(LOSS_VARIANTS[name])
    mod = yad_b200.AudioDetectionLoss(yad_b200.default_config()["anchors"], 2, **lc)
    tg = torch.from_numpy(g["targets"]).to(cuda_dev)
    preds = [torch.from_numpy(tr[f"pred{i}"]).to(cuda_dev).requires_grad_(True) for i in range(3)]
    with torch.enable_grad():
        loss, met = mod(preds, tg)
        loss.backward()
    np.testing.assert_allclose(float(loss), float(g[f"{name}_loss"]), rtol=5e-6)
    for i in range(3):
        np.testing.assert_allclose(preds[i].grad.cpu().numpy(), g[f"{name}_grad{i}"], atol=2e-7, rtol=2e-4)
    for k in ("conf_loss", "class_loss", "mean_ciou", "accuracy", "f1"):
        np.testing.assert_allclose(met[k], float(g[f"{name}_{k}"]), rtol=2e-5)
    # random case vs the oracle
    gen = torch.Generator().manual_seed(77)
    B, T = 8, 150
    rp = []
    for G in (120, 60, 30):
        p = torch.randn(B, G, 3, 5, generator=gen) * 2
        p[..., 3] = torch.rand(B, G, 3, generator=gen) * 60
        p[..., 4] = torch.rand(B, G, 3, generator=gen) * 20 + 0.05
        rp.append(p)
    t2 = torch.zeros(T, 4)
    t2[:, 0] = torch.randint(0, B, (T,), generator=gen).float()
    t2[:, 1] = torch.randint(0, 2, (T,), generator=gen).float()
    t2[::9, 1] = -100.0
    t2[:, 2] = torch.rand(T, generator=gen) * 60
    t2[:, 3] = torch.rand(T, generator=gen) * 12 + 0.2
    t2[7] = t2[6]
    with torch.enable_grad():
        ref_p = [p.clone().requires_grad_(True) for p in rp]
        ref_loss, _ = O.detection_loss(ref_p, t2, O.DEFAULT_CONFIG["anchors"], 2, **LOSS_VARIANTS[name])
        ref_loss.backward()
        cu_p = [p.to(cuda_dev).requires_grad_(True) for p in rp]
        loss2, _ = mod(cu_p, t2.to(cuda_dev))
        loss2.backward()
    np.testing.assert_allclose(float(loss2), float(ref_loss), rtol=5e-5)
    for a, b in zip(cu_p, ref_p):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.numpy(), atol=3e-7, rtol=5e-4)


@pytest.mark.parametrize("seed,B,T", [(0, 4, 23), (1, 16, 200), (2, 3, 0), (3, 32, 900)])
def test_detection_loss_vs_oracle(seed, B, T, cuda_dev):
    """Random predictions / targets (incl. duplicates, ignore labels, the no-target case) vs the CPU oracle."""
    gen = torch.Generator().manual_seed(seed)
    preds = []
    for G in (120, 60, 30):
        p = torch.randn(B, G, 3, 5, generator=gen)
        p[..., 3] = torch.rand(B, G, 3, generator=gen) * 60
        p[..., 4] = torch.rand(B, G, 3, generator=gen) * 20 + 0.05
        preds.append(p)
    tg = torch.zeros(T, 4)
    if T:
        tg[:, 0] = torch.randint(0, B, (T,), generator=gen).float()
        tg[:, 1] = torch.randint(0, 2, (T,), generator=gen).float()
        tg[::7, 1] = -100.0                                   # ignore_index rows
        tg[:, 2] = torch.rand(T, generator=gen) * 60
        tg[:, 3] = torch.rand(T, generator=gen) * 12 + 0.2
        if T > 10:
            tg[5] = tg[4]                                     # exact duplicate -> duplicate (b,g,a) keys
    with torch.enable_grad():
        ref_p = [p.clone().requires_grad_(True) for p in preds]
        ref_loss, ref_met = O.detection_loss(ref_p, tg, O.DEFAULT_CONFIG["anchors"], 2)
        if ref_loss.requires_grad:
            ref_loss.backward()
        cu_p = [p.to(cuda_dev).requires_grad_(True) for p in preds]
        loss, met = _loss_module()(cu_p, tg.to(cuda_dev))
        loss.backward()
    # the oracle (like the reference) reduces in fp32, the kernels in fp64: up to ~1e-5 relative on 10^5-term means
    np.testing.assert_allclose(float(loss), float(ref_loss), rtol=5e-5)
    for a, b in zip(cu_p, ref_p):
        rg = b.grad.numpy() if b.grad is not None else np.zeros(tuple(b.shape), np.float32)
        np.testing.assert_allclose(a.grad.cpu().numpy(), rg, atol=3e-7, rtol=5e-4)
    for k in ("mean_ciou", "conf_loss", "class_loss", "avg_pos_conf", "avg_neg_conf"):
        if ref_met[k] == ref_met[k]:
            np.testing.assert_allclose(met[k], ref_met[k], rtol=2e-5, atol=1e-7)
        else:
            assert met[k] != met[k]


def test_host_pipeline_equals_direct_call(models, cuda_dev):
    """run_host_batch (chunked, double-buffered H2D) returns exactly what model(x) + process_model_outputs returns."""
    m = models[("deploy", "bf16")]
    x = synth.synth_clips(5, 22050 * 6, seed=4000, silence_tail_every=3)
    seg_d, bidx_d = yad_b200.process_model_outputs(m(x.to(cuda_dev), combine_scales=True), 0.1, 0.2)
    seg_h, bidx_h = yad_b200.run_host_batch(m, x.pin_memory(), 0.1, 0.2, chunk=2)
    np.testing.assert_array_equal(bidx_h.numpy(), bidx_d.cpu().numpy())
    np.testing.assert_array_equal(seg_h.numpy(), seg_d.cpu().numpy())


def test_int16_pcm_ingest_is_bit_identical(models, cuda_dev):
    """16-bit PCM (the sample format of audio files): x / 32768 happens inside the frontend kernel, so mel
    power, x_spectral and the predictions are BIT-identical to feeding the reference's float tensor (torchaudio.load with
    normalize=True yields exactly int16 / 32768).  Covers the 16-byte cp.async path (L % 8 == 0) and the ragged fallback."""
    m = models[("deploy", "bf16")]
    for L in (22050 * 6, 22050 * 6 + 4, 22050 * 6 + 3):
        xi = (synth.synth_clips(3, L, seed=77, silence_tail_every=3).clamp(-1, 1) * 32767).round().to(torch.int16)
        xf = xi.float() / 32768.0
        ta, tb = {}, {}
        with torch.no_grad():
            pa = m(xi.to(cuda_dev), combine_scales=True, taps=ta)
            pb = m(xf.to(cuda_dev), combine_scales=True, taps=tb)
        assert torch.equal(ta["mel"], tb["mel"]), L
        assert torch.equal(ta["x_spectral"], tb["x_spectral"]), L
        assert torch.equal(pa, pb), L
    xi = (synth.synth_clips(5, 22050 * 6, seed=78).clamp(-1, 1) * 32767).round().to(torch.int16)
    sa = yad_b200.run_host_batch(m, xi.pin_memory(), 0.1, 0.05, chunk=2)
    sb = yad_b200.run_host_batch(m, (xi.float() / 32768.0).pin_memory(), 0.1, 0.05, chunk=2)
    assert sa[0] is not None and torch.equal(sa[0], sb[0]) and torch.equal(sa[1], sb[1])


def test_evaluate_waveform_vs_live_reference(models, gold, cuda_dev):
    """File-level chunker through the host-buffer pipeline == the live reference's evaluate_audio: same clips
    kept, same labels, times within 2e-3 s (f32 parity model), the same CSV rows; the int16 entry keeps the same segments."""
    from test_oracle_golden import _eval_expected, _rows_as_csv
    seg_e, bidx_e, rows_e = _eval_expected(gold)
    m = models[("train", "f32")]
    wav = synth.eval_waveform()
    seg, bidx, rows = yad_b200.evaluate_waveform(m, wav, synth.EVAL_SR, 60, 2, {0: "speech", 1: "music"}, synth.EVAL_IOU, synth.EVAL_CONF)
    np.testing.assert_array_equal(bidx.numpy(), bidx_e.numpy())
    np.testing.assert_array_equal(seg[:, 2].numpy(), seg_e[:, 2].numpy())
    np.testing.assert_allclose(seg.numpy(), seg_e.numpy(), atol=2e-3, rtol=1e-4)
    assert _rows_as_csv(rows) == rows_e
    wi = (wav.clamp(-1, 1) * 32767).round().to(torch.int16)
    seg16, bidx16, _ = yad_b200.evaluate_waveform(m, wi, synth.EVAL_SR, 60, 2, {0: "speech", 1: "music"}, synth.EVAL_IOU, synth.EVAL_CONF)
    assert seg16.shape == seg.shape and torch.equal(bidx16, bidx)        # 16-bit quantisation does not move a keep decision here


def test_evaluate_waveform_file_rate_vs_live_reference(models, gold, cuda_dev):
    """The chunker's file-rate branch (inference.py:152-159): a 16 kHz file through yad_resample_sinc + the 22.05 kHz model ==
    the live reference's evaluate_audio (same clips kept, same labels, times within 2e-3 s, same CSV rows); the resampling
    kernel alone == torchaudio.transforms.Resample for three rate pairs (fp32) and its int16 entry == the fp32 one on x / 32768."""
    from test_oracle_golden import _eval_expected, _rows_as_csv
    from yad_b200.evaluate import resample_to_model_rate
    g = gold("eval_rate16k")
    x = torch.from_numpy(g["rs_x"]).to(cuda_dev)
    for o, n in ((16000, 22050), (48000, 22050), (44100, 22050)):
        got = resample_to_model_rate(x, o, n)
        np.testing.assert_allclose(got.cpu().numpy(), g[f"rs_{o}_{n}"], atol=3e-6, rtol=1e-5)
    xi = (x.clamp(-1, 1) * 32767).round().to(torch.int16)
    np.testing.assert_array_equal(resample_to_model_rate(xi, 48000, 22050).cpu().numpy(),
                                  resample_to_model_rate(xi.float() / 32768.0, 48000, 22050).cpu().numpy())
    seg_e, bidx_e, rows_e = _eval_expected(gold, "eval_rate16k")
    m = models[("train", "f32")]
    seg, bidx, rows = yad_b200.evaluate_waveform(m, synth.eval_waveform_rate(), synth.EVAL_RATE2, 60, 2, {0: "speech", 1: "music"},
                                                 synth.EVAL_IOU, synth.EVAL_CONF)
    np.testing.assert_array_equal(bidx.numpy(), bidx_e.numpy())
    np.testing.assert_array_equal(seg[:, 2].numpy(), seg_e[:, 2].numpy())
    np.testing.assert_allclose(seg.numpy(), seg_e.numpy(), atol=2e-3, rtol=1e-4)
    assert _rows_as_csv(rows) == rows_e


# ------------------------------------------------------------------ other config-selectable backbones
@pytest.mark.parametrize("variant", ["bottleneck", "custom", "taper"])
@pytest.mark.parametrize("form", ["train", "deploy"])
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_other_backbones_vs_golden(variant_state_dict, gold, variant, form, dt, cuda_dev):
    """``resnet_config.block: Bottleneck`` and ``backbone: custom`` (3x7 ExtractorLayers, 2-D neck with H = 32) against the
    LIVE reference's eval() outputs (tests/golden/backbones.npz).  fp32 mode: the tolerance of the default net (5e-3);
    bf16 mode: the default net's stated bf16 tolerance."""
    g = gold("backbones")
    sd, cfg = variant_state_dict(variant)
    m = yad_b200.AudioDetectionNetwork(2, config=cfg, compute_dtype=dt)
    m.load_state_dict(sd)
    if form == "deploy":
        m.inference()
    m = m.eval().to(cuda_dev)
    x = synth.synth_clips(2, 22050 * 6, seed=3000, silence_tail_every=0).to(cuda_dev)
    taps = {}
    out = m(x, combine_scales=True, taps=taps).cpu().numpy()
    ref = g[f"{variant}.preds_{form}"]
    assert out.shape == ref.shape == (2, 63, 5)
    if dt == "f32":
        np.testing.assert_allclose(out, ref, atol=5e-3)
        if form == "train":
            np.testing.assert_allclose(taps["fmaps"][0][:, ::8, ::4, ::2].cpu().numpy(), g[f"{variant}.fmap1_s"], atol=2e-3)
            np.testing.assert_allclose(taps["fmaps"][3][:, ::8, ::4, :].cpu().numpy(), g[f"{variant}.fmap4_s"], atol=2e-3)
            for i, h in enumerate(taps["heads"]):
                np.testing.assert_allclose(h.cpu().numpy(), g[f"{variant}.head{i}"], atol=5e-3)
    else:
        d = np.abs(out - ref)
        assert d[..., :3].max() < 0.15 and d[..., :3].mean() < 0.02, (d[..., :3].max(), d[..., :3].mean())
        assert d[..., 3].max() < 0.1, d[..., 3].max()
        assert d[..., 4].max() < 1.5 and d[..., 4].mean() < 0.15, (d[..., 4].max(), d[..., 4].mean())
    sm, md, lg = m(x)
    assert sm.shape == (2, 12, 3, 5) and md.shape == (2, 6, 3, 5) and lg.shape == (2, 3, 3, 5)


# ------------------------------------------------------------------ anchor clustering
@pytest.mark.parametrize("name,init", [("a", "k-means++"), ("b", "k-means++"), ("c", "random")])
def test_kmeans_anchors_gpu(gold, name, init, cuda_dev):
    """yad_kmeans1d_lloyd vs the live sklearn run of compute_anchors.py (fixture) and the oracle: identical labels and
    iteration counts, centres to 1e-11 relative (fp64, different summation order)."""
    g = gold("anchors")
    d = g[f"{name}.durations"]
    sm, md, lg = yad_b200.compute_anchors(d, init=init, random_state=np.random.RandomState(42), device=cuda_dev)
    np.testing.assert_allclose(np.concatenate([sm, md, lg]), g[f"{name}.anchors"], rtol=1e-11)
    x = torch.from_numpy(d - d.mean()).to(cuda_dev)
    r = yad_b200.kmeans_lloyd(x, g[f"{name}.lloyd_init"] - d.mean(), 500, float(np.var(d)) * 1e-10)
    np.testing.assert_allclose(r["centers"] + d.mean(), g[f"{name}.lloyd_centers"], rtol=1e-11)
    np.testing.assert_array_equal(r["labels"].cpu().numpy(), g[f"{name}.lloyd_labels"])
    assert r["n_iter"] == int(g[f"{name}.lloyd_n_iter"])
    ol, oi, oc, on = O.kmeans_lloyd_1d(d - d.mean(), g[f"{name}.lloyd_init"] - d.mean(), 500, float(np.var(d)) * 1e-10)
    np.testing.assert_allclose(r["inertia"], oi, rtol=1e-11)


def test_kmeans_empty_cluster_gpu(cuda_dev):
    x = np.concatenate([np.linspace(-5, -4, 20), np.linspace(4, 5, 20), [30.0]])
    x = x - x.mean()
    c0 = np.array([x.min(), x.min() + 0.5, 1e3])
    r = yad_b200.kmeans_lloyd(torch.from_numpy(x).to(cuda_dev), c0, 100, 0.0)
    ol, oi, oc, on = O.kmeans_lloyd_1d(x, c0, 100, 0.0)
    np.testing.assert_allclose(np.sort(r["centers"]), np.sort(oc), rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(r["labels"].cpu().numpy(), ol)
    with pytest.raises(yad_b200.YadError):
        yad_b200.kmeans_lloyd(torch.zeros(4, dtype=torch.float64, device=cuda_dev), np.zeros(17), 10, 0.0)


# ------------------------------------------------------------------ checkpoint plumbing (pipeline/_trainer.py:38-53)
def test_checkpoint_roundtrip_with_torch_adam(tmp_path, cuda_dev):
    """A checkpoint in the reference's layout ({"network_params", "optimizer_params"} with torch.optim.Adam's state dict) resumes
    into FusedAdamEMA: the next fused step equals torch.optim.Adam's next step; and the other way round."""
    g = torch.Generator().manual_seed(5)
    w0 = [torch.randn(7, 5, generator=g), torch.randn(11, generator=g)]
    grads = [[torch.randn(7, 5, generator=g), torch.randn(11, generator=g)] for _ in range(4)]
    kw = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.002)
    ref_p = [torch.nn.Parameter(w.clone()) for w in w0]
    ref = torch.optim.Adam(ref_p, **kw)
    for s in range(2):
        for p, gr in zip(ref_p, grads[s]):
            p.grad = gr.clone()
        ref.step()
    mid = {"network_params": {f"p{i}": p.data.clone() for i, p in enumerate(ref_p)}, "optimizer_params": ref.state_dict()}
    torch.save(mid, tmp_path / "ck.pth.tar")
    saved = torch.load(tmp_path / "ck.pth.tar", map_location=cuda_dev)
    ps = [torch.nn.Parameter(saved["network_params"][f"p{i}"].clone()) for i in range(2)]
    opt = yad_b200.FusedAdamEMA(ps, lr=9.0)            # hyper-parameters come from the checkpoint
    opt.load_state_dict(saved["optimizer_params"])
    assert opt.step_count == 2 and opt.lr == 1e-3 and opt.wd == 0.002
    for s in (2, 3):
        for p, q, gr in zip(ref_p, ps, grads[s]):
            p.grad = gr.clone()
            q.grad.copy_(gr.to(cuda_dev))
        ref.step()
        opt.step()
    for p, q in zip(ref_p, ps):
        np.testing.assert_allclose(q.data.cpu().numpy(), p.data.numpy(), atol=2e-6)
    # and back: our state dict loads into torch.optim.Adam with the same moments
    sd = opt.state_dict()
    assert set(sd) == set(ref.state_dict()) and set(sd["param_groups"][0]) == set(ref.state_dict()["param_groups"][0])
    back = torch.optim.Adam([torch.nn.Parameter(q.data.cpu().clone()) for q in ps], **kw)
    back.load_state_dict({"state": {k: {kk: (vv.cpu() if torch.is_tensor(vv) else vv) for kk, vv in v.items()} for k, v in sd["state"].items()},
                          "param_groups": sd["param_groups"]})
    for i in range(2):
        np.testing.assert_allclose(back.state_dict()["state"][i]["exp_avg"].numpy(), ref.state_dict()["state"][i]["exp_avg"].numpy(), atol=1e-7)
        assert float(back.state_dict()["state"][i]["step"]) == 4.0


def test_save_load_checkpoint_model(tmp_path, models, cuda_dev):
    m = models[("train", "f32")]
    ps = [p for p in m.parameters()]
    x = synth.synth_clips(1, 22050 * 6, seed=77, silence_tail_every=0).to(cuda_dev)
    want = m(x, combine_scales=True).clone()

    class _Opt:
        def state_dict(self):
            return {"state": {}, "param_groups": []}
    path = str(tmp_path / "saved_model" / "AudioDetectionNetwork.pth.tar")
    yad_b200.save_checkpoint(path, m, _Opt())
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"network_params", "optimizer_params"} and len(ck["network_params"]) == 357
    m2 = yad_b200.AudioDetectionNetwork(2, compute_dtype="f32").eval().to(cuda_dev)
    yad_b200.load_checkpoint(path, m2, device=cuda_dev)
    np.testing.assert_array_equal(m2(x, combine_scales=True).cpu().numpy(), want.cpu().numpy())
    with pytest.raises(OSError):
        yad_b200.load_checkpoint(str(tmp_path / "nope.pth.tar"), m2)
    assert len(ps) == 182


# ------------------------------------------------------------------ fused stem: conv1 o conv2 as one 19x19 stride-4 convolution
@pytest.mark.parametrize("B,T", [(2, 96), (1, 960), (3, 61), (2, 130)])
def test_fused_stem_vs_two_convs(B, T, cuda_dev):
    """yad_conv_stem_fused + fix-up against F.conv2d(F.conv2d(x, W1), W2 * bn) in fp32 on the bf16-rounded input (the kernel
    rounds x and the composite weights to bf16, accumulates in fp32): every output pixel incl. the border rows / columns whose
    conv2 taps leave conv1's output (zero padding of the intermediate tensor)."""
    from yad_b200.engine import InferenceEngine
    g = torch.Generator().manual_seed(B * 1000 + T)
    w1 = (torch.rand(64, 2, 7, 7, generator=g) * 2 - 1) * 0.17
    w2 = (torch.rand(64, 64, 7, 7, generator=g) * 2 - 1) * 0.03
    b2 = torch.randn(64, generator=g) * 0.2
    x = torch.randn(B, 2, 32, T, generator=g)
    eng = InferenceEngine.__new__(InferenceEngine)
    eng.dev = cuda_dev
    eng._pack_fused_stem(w1, w2, b2)
    lib = _lib.init(0)
    H1, W1 = 16, (T - 1) // 2 + 1
    Ho, Wo = 8, (W1 - 1) // 2 + 1
    Hp, Wp = Ho + 1, Wo + 1
    out = torch.zeros(B, Wp, Hp, 64, device=cuda_dev, dtype=torch.bfloat16)
    xd = x.to(cuda_dev)
    # padded channel-interleaved bf16 words (what yad_frontend_finish_bf16 writes): 9-word zero margin, zero tail
    n_seg = (Wo + 127) // 128
    pitch = (max(512 * (n_seg - 1) + 536, 9 + T) + 3) // 4 * 4
    xb16 = torch.zeros(B, 32, pitch, 2, device=cuda_dev, dtype=torch.bfloat16)
    xb16[:, :, 9:9 + T, :] = xd.permute(0, 2, 3, 1).to(torch.bfloat16)
    _lib.check(lib.yad_conv_stem_fused(xb16.data_ptr(), pitch, B, 32, T, eng.fstem_w.data_ptr(), eng.fstem_bias.data_ptr(), out.data_ptr(), Hp,
                                       Wp, 0, _stream()), "fused")
    cols, var = InferenceEngine._fused_stem_border_cols(T)
    _lib.check(lib.yad_conv_stem_fused_fixup(xd.data_ptr(), B, 32, T, eng.fstem_wvar.data_ptr(), eng.fstem_bias.data_ptr(),
                                             (C.c_int32 * len(cols))(*cols), (C.c_int32 * len(var))(*var), len(cols), out.data_ptr(), Hp, Wp,
                                             _stream()), "fixup")
    got = out[:, :Wo, :Ho, :].float().permute(0, 3, 2, 1).cpu()            # [B, 64, Ho, Wo]
    xb = x.bfloat16().float()
    ref = F.relu(F.conv2d(F.conv2d(xb.double(), w1.double(), None, 2, 3), w2.double(), b2.double(), 2, 3)).float()
    assert got.shape == ref.shape
    scale = ref.abs().max().item()
    err = (got - ref).abs()
    assert err.max().item() < 2e-2 * scale, (err.max().item(), scale)      # bf16 weights / output rounding
    assert err.mean().item() < 2e-3 * scale
    # the halo cells stay zero
    assert out[:, Wo:, :, :].abs().max().item() == 0 and out[:, :, Ho:, :].abs().max().item() == 0


def test_recorded_call_list_replay(models, cuda_dev):
    """The first forward of a (batch, length) plan records its C-ABI calls; later forwards replay them with the input pointer,
    the output and the stream patched in.  Replays must equal a fresh engine's first run bit for bit, also on another tensor
    and on another stream."""
    m = models[("deploy", "bf16")]
    xa = synth.synth_clips(2, 22050 * 6, seed=501, silence_tail_every=0).to(cuda_dev)
    xb = synth.synth_clips(2, 22050 * 6, seed=502, silence_tail_every=0).to(cuda_dev)
    m._engine_cache.clear()
    first_a = m(xa, combine_scales=True).clone()            # recorded run
    eng = m._engine()
    assert any(isinstance(k, tuple) and k[0] == "prog" for k in eng._plan((2, 22050 * 6)))
    rep_a = m(xa, combine_scales=True).clone()              # replay, same tensor
    rep_b = m(xb, combine_scales=True).clone()              # replay, other tensor
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        rep_a2 = m(xa, combine_scales=True).clone()         # replay on another stream
    side.synchronize()
    m._engine_cache.clear()
    first_b = m(xb, combine_scales=True).clone()            # fresh engine, recorded run on the other tensor
    torch.cuda.synchronize()
    assert torch.equal(first_a, rep_a) and torch.equal(first_a, rep_a2) and torch.equal(first_b, rep_b)
    assert not torch.equal(first_a, first_b)


@pytest.mark.parametrize("seconds", [2.0, 10.0, 22.0])
def test_fused_stem_network_equals_two_conv_path(models, seconds, cuda_dev, monkeypatch):
    """Whole bf16 network with the composed stem against the same network with conv1 and conv2 as separate tensor-core
    convolutions (YAD_FUSED_STEM=0) at other clip lengths (T = 32, 160, 352 frames: one and two column segments, the shortest
    input the net takes; the odd-width border variants are covered by test_fused_stem_vs_two_convs).  The two differ only by
    bf16 rounding of the intermediate tensor / of the composite weights."""
    m = models[("deploy", "bf16")]
    L = int(22050 * seconds) // 4 * 4
    x = synth.synth_clips(3, L, seed=900 + int(seconds), silence_tail_every=0).to(cuda_dev)
    m._engine_cache.clear()
    fused = m(x, combine_scales=True).clone()
    assert m._engine().fused_stem
    monkeypatch.setenv("YAD_FUSED_STEM", "0")
    m._engine_cache.clear()
    plain = m(x, combine_scales=True).clone()
    assert not m._engine().fused_stem
    monkeypatch.delenv("YAD_FUSED_STEM")
    m._engine_cache.clear()
    d = (fused - plain).abs()
    assert d[..., :3].max().item() < 0.1 and d[..., :3].mean().item() < 0.01, (d[..., :3].max().item(), d[..., :3].mean().item())
    assert d[..., 3].max().item() < 0.1 and d[..., 4].max().item() < 1.0


@pytest.mark.parametrize("seconds,B", [(2.0, 3), (10.0, 2), (22.0, 5), (60.0, 3), (120.0, 2)])
def test_fused_neck_equals_layer_by_layer_neck(models, seconds, B, cuda_dev, monkeypatch):
    """The one-kernel neck (yad_neck_fused: H-means folded into K, every intermediate in shared memory; reference
    modules/_common.py:241-265) against the layer-by-layer neck (YAD_FUSED_NECK=0: yad_hmean + 21 convolutions + glue kernels)
    on the same bf16 deploy-form network, at clip lengths giving 1..2 M tiles per level (W1 = 8, 40, 88, 240).  They differ by
    the bf16 rounding of the pooled maps only (the fused kernel accumulates the H-mean in fp32 inside the GEMM)."""
    m = models[("deploy", "bf16")]
    L = int(22050 * seconds) // 4 * 4
    x = synth.synth_clips(B, L, seed=1300 + int(seconds), silence_tail_every=2).to(cuda_dev)
    m._engine_cache.clear()
    fused = m(x, combine_scales=True).clone()
    eng = m._engine()
    assert eng.fused_neck and any(v is not None for v in eng.__dict__.get("_fused_necks", {}).values())
    monkeypatch.setenv("YAD_FUSED_NECK", "0")
    m._engine_cache.clear()
    plain = m(x, combine_scales=True).clone()
    assert not m._engine().fused_neck
    monkeypatch.delenv("YAD_FUSED_NECK")
    m._engine_cache.clear()
    assert torch.isfinite(fused).all()
    d = (fused - plain).abs()
    assert d[..., :3].max().item() < 0.1 and d[..., :3].mean().item() < 0.01, (d[..., :3].max().item(), d[..., :3].mean().item())
    assert d[..., 3].max().item() < 0.1 and d[..., 4].max().item() < 1.0


@pytest.mark.parametrize("seconds,B", [(2.0, 3), (22.0, 5), (60.0, 3), (60.0, 8)])
def test_fused_neck_two_clips_per_pass_equals_one(models, seconds, B, cuda_dev, monkeypatch):
    """Fused neck with two clips per CTA pass (the default: clips of a unit at a power-of-two row pitch, both in one M tile at
    levels 3 / 4, one weight fetch per pair, the level-1 conv once per clip; csrc/neck_fused.cu) against one clip per pass
    (YAD_NECK_G=1: the backbone's own W + 1 row pitch).  Same program, same K order per output row: BITWISE equal - including odd
    batch sizes (the last unit holds one clip and one out-of-range clip) and clips whose halo / padding rows must read back zero
    for the neighbouring clip's 3-tap convolutions (modules/_common.py:241-265)."""
    m = models[("deploy", "bf16")]
    L = int(22050 * seconds) // 4 * 4
    x = synth.synth_clips(B, L, seed=1700 + int(seconds) + B, silence_tail_every=2).to(cuda_dev)
    m._engine_cache.clear()
    two = m(x, combine_scales=True).clone()
    necks = [v for v in m._engine().__dict__.get("_fused_necks", {}).values() if v is not None]
    assert necks and necks[0].G == 2
    monkeypatch.setenv("YAD_NECK_G", "1")
    m._engine_cache.clear()
    one = m(x, combine_scales=True).clone()
    necks = [v for v in m._engine().__dict__.get("_fused_necks", {}).values() if v is not None]
    assert necks and necks[0].G == 1
    monkeypatch.delenv("YAD_NECK_G")
    m._engine_cache.clear()
    assert torch.isfinite(two).all()
    assert torch.equal(two, one), float((two - one).abs().max())


def _run_with_env(m, x, monkeypatch, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    m._engine_cache.clear()
    taps = {}
    out = m(x, combine_scales=True, taps=taps).clone()
    eng = m._engine()
    for k in env:
        monkeypatch.delenv(k)
    m._engine_cache.clear()
    return out, taps["fmaps"], eng


@pytest.mark.parametrize("seconds,B", [(2.0, 3), (22.0, 2), (60.0, 2)])
def test_stride2_block_routes_agree(models, seconds, B, cuda_dev, monkeypatch):
    """First BasicBlock of layer2..4 (torchvision resnet.py:92-100 via modules/_backbone.py:148): conv1 (3x3 stride 2) and the
    1x1 stride-2 downsample of the same input, three ways:
      a) two tap-by-tap launches (yad_conv_tc; YAD_S2D=0 YAD_DUAL_DS=0),
      b) one tap-by-tap launch with two TMEM accumulators (yad_conv_tc_dual; YAD_S2D=0): same tiles and K order -> bitwise a),
      c) default: the previous layer also writes a space-to-depth copy (yad_conv_flat_s2d); conv1 runs in the patch-resident
         kernel over (plane, shift) steps (yad_conv_flat_taps) and the downsample rides in conv2's GEMM as extra K steps
         (yad_conv_flat_taps2): other K order, identity not rounded to bf16 -> equal up to fp32 summation order and bf16 rounding
         (checked on the backbone maps and on the predictions)."""
    m = models[("deploy", "bf16")]
    L = int(22050 * seconds) // 4 * 4
    x = synth.synth_clips(B, L, seed=1500 + int(seconds), silence_tail_every=0).to(cuda_dev)
    out_a, fm_a, eng_a = _run_with_env(m, x, monkeypatch, {"YAD_S2D": "0", "YAD_DUAL_DS": "0"})
    assert not eng_a.dual_ds and not eng_a.s2d_route
    out_b, fm_b, eng_b = _run_with_env(m, x, monkeypatch, {"YAD_S2D": "0"})
    assert eng_b.dual_ds and not eng_b.s2d_route
    for a, b in zip(fm_a, fm_b):
        assert torch.equal(a, b)
    assert torch.equal(out_a, out_b)
    out_c, fm_c, eng_c = _run_with_env(m, x, monkeypatch, {})
    assert eng_c.s2d_route
    for i, (a, c) in enumerate(zip(fm_a, fm_c)):
        assert a.shape == c.shape and torch.isfinite(c).all()
        d = (a - c).abs()
        scale = a.abs().max().item()
        assert d.max().item() <= 0.02 * scale + 1e-3 and d.mean().item() <= 5e-4 * scale, (i, d.max().item(), d.mean().item(), scale)
    assert torch.equal(fm_a[0], fm_c[0])          # layer1 is untouched by the route
    d = (out_a - out_c).abs()
    # same tolerance as the other bf16 route comparisons (test_fused_stem_network_equals_two_conv_path)
    assert d[..., :3].max().item() < 0.1 and d[..., :3].mean().item() < 0.01, (d[..., :3].max().item(), d[..., :3].mean().item())
    assert d[..., 3].max().item() < 0.1 and d[..., 4].max().item() < 1.0


# ------------------------------------------------------------------ small kernels added for the 2-D neck / fused stem
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_maxpool_h_and_2d_sppf_vs_torch(dt, cuda_dev):
    """yad_sppf_pools (W cascades, B*H rows) followed by yad_maxpool_h (box maxima over 5 / 9 / 13 rows) = three cascaded
    F.max_pool2d(k=5, s=1, p=2) of the reference's CSPSPPFModule when the neck keeps its height (modules/_common.py:197,207-209)."""
    lib = _lib.init(0)
    tdt, code = (torch.float32, F32) if dt == "f32" else (torch.bfloat16, BF16)
    B, H, W, Cc = 2, 32, 7, 64
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, H, W, Cc, generator=g).to(tdt).to(cuda_dev)
    wp = torch.zeros(B, H, W, 192, device=cuda_dev, dtype=tdt)
    out = torch.zeros(B, H, W, 256, device=cuda_dev, dtype=tdt)
    out[..., :64] = x
    _lib.check(lib.yad_sppf_pools(out.data_ptr(), code, B * H, W, 64, 256, 0, wp.data_ptr(), 192, 0, _stream()), "sppf")
    for k in range(3):
        _lib.check(lib.yad_maxpool_h(wp.data_ptr(), code, B, H, W, 64, 192, 64 * k, 2 * (k + 1), out.data_ptr(), 256, 64 * (k + 1), _stream()), "mph")
    xn = x.float().permute(0, 3, 1, 2)
    p1 = F.max_pool2d(xn, 5, 1, 2); p2 = F.max_pool2d(p1, 5, 1, 2); p3 = F.max_pool2d(p2, 5, 1, 2)
    ref = torch.cat([xn, p1, p2, p3], 1).permute(0, 2, 3, 1)
    assert torch.equal(out.float(), ref)             # maxima of the same values: exact in both dtypes


def test_nchw_to_nhwc_and_bf16_stage_b_copy(models, cuda_dev):
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(3, 2, 32, 50, generator=g).to(cuda_dev)
    for tdt, code in ((torch.float32, F32), (torch.bfloat16, BF16)):
        o = torch.zeros(3, 32, 50, 2, device=cuda_dev, dtype=tdt)
        _lib.check(lib.yad_nchw_to_nhwc(x.data_ptr(), 3, 2, 32, 50, o.data_ptr(), code, 2, _stream()), "nchw_to_nhwc")
        assert torch.equal(o, x.permute(0, 2, 3, 1).to(tdt))
    # stage B's padded channel-interleaved bf16 copy = x_spectral rounded to bf16, margins untouched (zero)
    m = models[("deploy", "bf16")]
    clips = synth.synth_clips(2, 22050 * 6, seed=77, silence_tail_every=0).to(cuda_dev)
    m._engine_cache.clear()
    m(clips, combine_scales=True)
    eng = m._engine()
    plan = eng._plan((2, 22050 * 6))
    xs, xb = plan["xs"], plan["xs_bf16"]
    T = xs.shape[-1]
    words = xb.view(torch.bfloat16).reshape(2, 32, xb.shape[2], 2)
    assert torch.equal(words[:, :, 9:9 + T, :], xs.permute(0, 2, 3, 1).to(torch.bfloat16))
    assert words[:, :, :9].abs().max().item() == 0 and words[:, :, 9 + T:].abs().max().item() == 0


@pytest.mark.parametrize("iou,cthr", [(0.1, 0.2), (0.05, 0.5), (0.1, 0.9999)])
def test_nms_segments_only_equals_full_scan(gold, iou, cthr, cuda_dev):
    """want_keep=False lets the greedy scan stop at the confidence threshold: the segments must be identical to the full scan's
    (a box can only be suppressed by a higher-scored one, and the reference filters after the NMS: inference.py:75-88)."""
    for seed, B in ((5, 4), (9, 2)):
        o = synth.synth_heads(B, 630, 2, seed=seed).to(cuda_dev)
        full = yad_b200.nms_raw(o, iou, cthr)
        fast = yad_b200.nms_raw(o, iou, cthr, want_keep=False)
        assert "keep" not in fast
        assert torch.equal(full["n_seg"], fast["n_seg"])
        for b in range(B):
            n = int(full["n_seg"][b])
            assert torch.equal(full["seg_rows"][b, :n], fast["seg_rows"][b, :n])


def test_graph_replay_of_repeated_input(models, cuda_dev):
    """The same input TENSOR coming back turns the recorded plan into one CUDA-graph launch: identical results, it reads the
    tensor's current contents, returns fresh output tensors, and a different tensor falls back to the plain replay."""
    m = models[("deploy", "bf16")]
    xa = synth.synth_clips(2, 22050 * 6, seed=611, silence_tail_every=0).to(cuda_dev)
    xb = synth.synth_clips(2, 22050 * 6, seed=612, silence_tail_every=0).to(cuda_dev)
    m._engine_cache.clear()
    outs = [m(xa, combine_scales=True) for _ in range(5)]             # record, replay, capture + graph, graph, graph
    eng = m._engine()
    progs = [v for k, v in eng._plan((2, 22050 * 6)).items() if isinstance(k, tuple) and k[0] == "prog"]
    assert progs and len(progs[0].get("graphs", {})) == 1
    for o in outs[1:]:
        assert torch.equal(o, outs[0]) and o.data_ptr() != outs[0].data_ptr()
    ref_b = m(xb, combine_scales=True).clone()                        # other tensor, first sighting: plain replay
    assert not torch.equal(ref_b, outs[0])
    assert len(progs[0]["graphs"]) == 1
    for _ in range(2):                                                # it comes back: its own graph (a ring of staging buffers)
        assert torch.equal(m(xb, combine_scales=True), ref_b)
    assert len(progs[0]["graphs"]) == 2
    xa.copy_(xb)                                                      # same tensor, new contents: the graph reads them
    assert torch.equal(m(xa, combine_scales=True), ref_b)


def test_shared_model_across_threads_with_graph_replay(models, cuda_dev):
    """inference.py:212-236 hands ONE model to 10 worker threads.  Each thread has its own plans (workspaces, recorded call lists,
    CUDA graphs): four threads calling the model repeatedly with their own tensors, while graphs are being captured, must each
    get the single-threaded result."""
    import threading
    m = models[("deploy", "bf16")]
    m._engine_cache.clear()
    xs_ = [synth.synth_clips(2, 22050 * 6, seed=700 + i, silence_tail_every=0).to(cuda_dev) for i in range(4)]
    want = [m(x, combine_scales=True).clone() for x in xs_]
    torch.cuda.synchronize()
    got, errs = [None] * 4, []

    def work(i):
        try:
            torch.cuda.set_device(cuda_dev)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(5):
                    o = m(xs_[i], combine_scales=True)
                st.synchronize()
            got[i] = o.clone()
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i in range(4):
        assert torch.equal(got[i], want[i]), i
