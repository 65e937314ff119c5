"""Train-mode kernels against PyTorch fp32 autograd of the same ops (run with -m gpu on a B200)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import yad_b200  # noqa: F401
from oracle import ref_port as O
from yad_b200 import _lib
from yad_b200._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ConvDesc

pytestmark = pytest.mark.gpu
# the PyTorch reference must really be fp32: cuDNN convolutions default to TF32 (10-bit mantissa) on this GPU
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(autouse=True)
def _grad_enabled():
    """Other test modules switch autograd off globally; these tests compare against PyTorch autograd."""
    with torch.enable_grad():
        yield


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _act(y, act):
    return F.relu(y) if act == ACT_RELU else (F.leaky_relu(y, 0.2) if act == ACT_LRELU else y)


CONV_CASES = [
    # B, H, W, Cin, Cout, (kh, kw), (sh, sw), (ph, pw)
    (2, 8, 24, 64, 64, (3, 3), (1, 1), (1, 1)),
    (2, 8, 24, 64, 128, (3, 3), (2, 2), (1, 1)),
    (2, 8, 24, 64, 128, (1, 1), (2, 2), (0, 0)),
    (2, 16, 48, 64, 64, (7, 7), (2, 2), (3, 3)),
    (3, 1, 24, 15, 128, (3, 3), (1, 2), (1, 1)),
    (2, 32, 40, 2, 64, (7, 7), (2, 2), (3, 3)),
    (3, 1, 12, 128, 15, (3, 3), (1, 1), (1, 1)),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_backward(case, cuda_dev):
    B, H, W, Cin, Cout, k, s, p = case
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(abs(hash(case)) % 9973)
    x = torch.randn(B, Cin, H, W, generator=g).to(cuda_dev).requires_grad_(True)
    w = (torch.randn(Cout, Cin, *k, generator=g) / (Cin * k[0] * k[1]) ** 0.5).to(cuda_dev).requires_grad_(True)
    b = torch.randn(Cout, generator=g).to(cuda_dev).requires_grad_(True)
    y = F.conv2d(x, w, b, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g).to(cuda_dev)
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    d = ConvDesc(B=B, H=H, W=W, Cin=Cin, ld_in=Cin, Cout=Cout, ld_out=Cout + 8, co_off=8, kh=k[0], kw=k[1], sh=s[0], sw=s[1],
                 ph=p[0], pw=p[1], act=ACT_NONE, ld_res=0)
    xh = x.detach().permute(0, 2, 3, 1).contiguous()
    dyh = torch.zeros(B, Ho, Wo, Cout + 8, device=cuda_dev)
    dyh[..., 8:] = dy.permute(0, 2, 3, 1)
    wt = w.detach().permute(2, 3, 0, 1).contiguous()                  # [kh][kw][Cout][Cin]
    prev = torch.randn(B, H, W, Cin, generator=g).to(cuda_dev)        # an existing gradient to accumulate onto
    dx = torch.empty_like(prev)
    _lib.check(lib.yad_conv_dgrad(C.byref(d), dyh.data_ptr(), wt.data_ptr(), prev.data_ptr(), dx.data_ptr(), _stream()), "dgrad")
    dw = torch.zeros(k[0], k[1], Cin, Cout, device=cuda_dev)
    db = torch.zeros(Cout, device=cuda_dev, dtype=torch.float64)
    _lib.check(lib.yad_conv_wgrad(C.byref(d), xh.data_ptr(), dyh.data_ptr(), dw.data_ptr(), db.data_ptr(), _stream()), "wgrad")
    torch.cuda.synchronize()
    np.testing.assert_allclose((dx - prev).permute(0, 3, 1, 2).cpu().numpy(), x.grad.cpu().numpy(), atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(dw.permute(3, 2, 0, 1).cpu().numpy(), w.grad.cpu().numpy(), atol=2e-3, rtol=1e-4)
    np.testing.assert_allclose(db.cpu().numpy(), b.grad.double().cpu().numpy(), atol=1e-4, rtol=1e-5)


@pytest.mark.parametrize("act", [ACT_NONE, ACT_RELU, ACT_LRELU])
@pytest.mark.parametrize("N,Cc", [(2 * 8 * 24, 64), (5000, 15), (37, 128)])
def test_batchnorm_train_fwd_bwd(act, N, Cc, cuda_dev):
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(N + Cc + act)
    x = (torch.randn(N, Cc, generator=g) * 2 + 0.5).to(cuda_dev).requires_grad_(True)
    gamma = (torch.rand(Cc, generator=g) + 0.5).to(cuda_dev).requires_grad_(True)
    beta = torch.randn(Cc, generator=g).to(cuda_dev).requires_grad_(True)
    rm, rv = torch.randn(Cc, generator=g).to(cuda_dev), (torch.rand(Cc, generator=g) + 0.5).to(cuda_dev)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y_ref = _act(F.batch_norm(x, rm_ref, rv_ref, gamma, beta, training=True, momentum=0.1, eps=1e-5), act)
    dy = torch.randn(N, Cc, generator=g).to(cuda_dev)
    y_ref.backward(dy)
    ld = Cc + 4
    xb = torch.zeros(N, ld, device=cuda_dev); xb[:, :Cc] = x.detach()
    y = torch.zeros(N, ld, device=cuda_dev)
    sm, si = torch.empty(Cc, device=cuda_dev), torch.empty(Cc, device=cuda_dev)
    ws = torch.empty(2 * Cc, device=cuda_dev, dtype=torch.float64)
    _lib.check(lib.yad_bn_train_fwd(xb.data_ptr(), ld, N, Cc, gamma.data_ptr(), beta.data_ptr(), 1e-5, 0.1, rm.data_ptr(), rv.data_ptr(),
                                    act, y.data_ptr(), ld, sm.data_ptr(), si.data_ptr(), ws.data_ptr(), _stream()), "bn fwd")
    dx = torch.zeros(N, ld, device=cuda_dev)
    dg, dbt = torch.ones(Cc, device=cuda_dev), torch.ones(Cc, device=cuda_dev)        # accumulate onto ones
    dyb = torch.zeros(N, ld, device=cuda_dev); dyb[:, :Cc] = dy
    _lib.check(lib.yad_bn_train_bwd(xb.data_ptr(), ld, y.data_ptr(), ld, dyb.data_ptr(), ld, N, Cc, gamma.data_ptr(), sm.data_ptr(),
                                    si.data_ptr(), act, dx.data_ptr(), ld, 0, dg.data_ptr(), dbt.data_ptr(), ws.data_ptr(), _stream()), "bn bwd")
    torch.cuda.synchronize()
    np.testing.assert_allclose(y[:, :Cc].cpu().numpy(), y_ref.detach().cpu().numpy(), atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(rm.cpu().numpy(), rm_ref.cpu().numpy(), atol=1e-6, rtol=1e-5)
    np.testing.assert_allclose(rv.cpu().numpy(), rv_ref.cpu().numpy(), atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(dx[:, :Cc].cpu().numpy(), x.grad.cpu().numpy(), atol=2e-5, rtol=2e-4)
    np.testing.assert_allclose((dg - 1).cpu().numpy(), gamma.grad.cpu().numpy(), atol=2e-3, rtol=2e-4)
    np.testing.assert_allclose((dbt - 1).cpu().numpy(), beta.grad.cpu().numpy(), atol=2e-3, rtol=2e-4)


def test_add_act_and_glue_backward(cuda_dev):
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(5)
    B, H, W, Cc = 3, 4, 12, 64
    # act(a + b + c)
    a, b, c = [torch.randn(B * W, Cc, generator=g).to(cuda_dev).requires_grad_(True) for _ in range(3)]
    y_ref = F.leaky_relu(a + b + c, 0.2)
    dy = torch.randn(B * W, Cc, generator=g).to(cuda_dev)
    y_ref.backward(dy)
    y = torch.empty(B * W, Cc, device=cuda_dev)
    _lib.check(lib.yad_add_act(a.data_ptr(), Cc, b.data_ptr(), Cc, c.data_ptr(), Cc, B * W, Cc, ACT_LRELU, y.data_ptr(), Cc, _stream()), "add")
    da, db, dc = [torch.zeros(B * W, Cc, device=cuda_dev) for _ in range(3)]
    _lib.check(lib.yad_add_act_bwd(y.data_ptr(), Cc, dy.data_ptr(), Cc, B * W, Cc, ACT_LRELU, da.data_ptr(), Cc, db.data_ptr(), Cc,
                                   dc.data_ptr(), Cc, _stream()), "add bwd")
    np.testing.assert_allclose(y.cpu().numpy(), y_ref.detach().cpu().numpy(), atol=1e-6)
    for got, ref in ((da, a), (db, b), (dc, c)):
        np.testing.assert_allclose(got.cpu().numpy(), ref.grad.cpu().numpy(), atol=1e-6)
    # H-mean
    x = torch.randn(B, Cc, H, W, generator=g).to(cuda_dev).requires_grad_(True)
    F.adaptive_avg_pool2d(x, (1, W)).backward(dyo := torch.randn(B, Cc, 1, W, generator=g).to(cuda_dev))
    din = torch.zeros(B, H, W, Cc, device=cuda_dev)
    do = dyo.permute(0, 2, 3, 1).contiguous()
    _lib.check(lib.yad_hmean_bwd(do.data_ptr(), Cc, B, H, W, Cc, din.data_ptr(), Cc, _stream()), "hmean bwd")
    np.testing.assert_allclose(din.permute(0, 3, 1, 2).cpu().numpy(), x.grad.cpu().numpy(), atol=1e-6)
    # bilinear x2 / x0.5 along W
    for up, sf in ((1, (1, 2)), (0, (1, 0.5))):
        x = torch.randn(B, Cc, 1, W, generator=g).to(cuda_dev).requires_grad_(True)
        yr = F.interpolate(x, scale_factor=sf, mode="bilinear")
        dyo = torch.randn(yr.shape, generator=g).to(cuda_dev)
        yr.backward(dyo)
        din = torch.zeros(B, 1, W, Cc, device=cuda_dev)
        do = dyo.permute(0, 2, 3, 1).contiguous()
        _lib.check(lib.yad_resize_w_bwd(do.data_ptr(), Cc, B, W, Cc, up, din.data_ptr(), Cc, _stream()), "resize bwd")
        np.testing.assert_allclose(din.permute(0, 3, 1, 2).cpu().numpy(), x.grad.cpu().numpy(), atol=1e-6)
    # MaxPool(5,1,2) along W
    x = torch.randn(B, Cc, 1, W, generator=g).to(cuda_dev).requires_grad_(True)
    yr = F.max_pool2d(x, kernel_size=5, stride=1, padding=2)
    dyo = torch.randn(yr.shape, generator=g).to(cuda_dev)
    yr.backward(dyo)
    xh = x.detach().permute(0, 2, 3, 1).contiguous()
    yh = torch.empty_like(xh)
    _lib.check(lib.yad_maxpool5_w(xh.data_ptr(), Cc, B, W, Cc, yh.data_ptr(), Cc, _stream()), "maxpool")
    dxh = torch.zeros_like(xh)
    do = dyo.permute(0, 2, 3, 1).contiguous()
    _lib.check(lib.yad_maxpool5_w_bwd(xh.data_ptr(), Cc, do.data_ptr(), Cc, B, W, Cc, dxh.data_ptr(), Cc, _stream()), "maxpool bwd")
    np.testing.assert_allclose(yh.permute(0, 3, 1, 2).cpu().numpy(), yr.detach().cpu().numpy(), atol=0)
    np.testing.assert_allclose(dxh.permute(0, 3, 1, 2).cpu().numpy(), x.grad.cpu().numpy(), atol=1e-6)


def test_dropout_mask(cuda_dev):
    lib = _lib.init(0)
    n = 1 << 20
    x = torch.ones(n, device=cuda_dev)
    y = torch.empty(n, device=cuda_dev)
    _lib.check(lib.yad_dropout(x.data_ptr(), n, 0.4, 1234, 0, y.data_ptr(), _stream()), "dropout")
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.6) < 3e-3 and torch.all((y == 0) | ((y - 1 / 0.6).abs() < 1e-6))
    y2 = torch.zeros(n, device=cuda_dev)
    _lib.check(lib.yad_dropout(x.data_ptr(), n, 0.4, 1234, 1, y2.data_ptr(), _stream()), "dropout")     # same seed -> same mask
    assert torch.equal(y, y2)
    _lib.check(lib.yad_dropout(x.data_ptr(), n, 0.0, 7, 0, y.data_ptr(), _stream()), "dropout p=0")
    assert torch.all(y == 1)


def test_decode_backward(cuda_dev):
    lib = _lib.init(0)
    g = torch.Generator().manual_seed(9)
    B, G, A, nc = 3, 12, 3, 2
    E = 3 + nc
    with torch.enable_grad():
        head = (torch.randn(B, G, A * E, generator=g) * 2).requires_grad_(True)
        anc = torch.tensor([0.02, 0.05, 0.4]).requires_grad_(True)
        pred = O.decode_scale(head, anc * 60.0, 96000, 96, nc)          # 6 s clips: 96 frames
        dp = torch.randn(pred.shape, generator=g)
        pred.backward(dp)
    ld = 16
    hb = torch.zeros(B, G, ld, device=cuda_dev); hb[..., : A * E] = head.detach().to(cuda_dev)
    dh = torch.zeros(B, G, ld, device=cuda_dev)
    da = torch.zeros(A, device=cuda_dev)
    anc_s = (anc.detach() * 60.0).to(cuda_dev)
    stride_over_scaler = (96 // G) / (96 / (96000 / 16000))
    _lib.check(lib.yad_decode_bwd(hb.data_ptr(), ld, dp.to(cuda_dev).contiguous().data_ptr(), B, G, A, nc, anc_s.data_ptr(),
                                  stride_over_scaler, 60.0, dh.data_ptr(), ld, da.data_ptr(), _stream()), "decode bwd")
    np.testing.assert_allclose(dh[..., : A * E].cpu().numpy(), head.grad.numpy(), atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose((da * 60.0).cpu().numpy(), anc.grad.numpy(), atol=1e-4, rtol=1e-4)


# ------------------------------------------------------------------ the train-mode network (rows a1 / a17)
def _train_model(ref_state_dict, cuda_dev, dropout=0.0, train_dtype="f32"):
    cfg = yad_b200.default_config()
    cfg["dropout"] = dropout
    m = yad_b200.AudioDetectionNetwork(2, config=cfg, train_dtype=train_dtype)
    m.load_state_dict(ref_state_dict)
    return m.to(cuda_dev).train()


def _loss_fn():
    cfg = yad_b200.default_config()
    return yad_b200.AudioDetectionLoss(cfg["anchors"], 2, sample_duration=60, **cfg["train_config"]["loss_config"])


def _rel_l2(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / max(float(b.norm()), 1e-12))


def test_train_network_forward_backward_vs_oracle_and_reference(gold, ref_state_dict, cuda_dev):
    """Train-mode CNN + decode on the oracle's own x_spectral (so that the ill-conditioned MFCC-dB channel of the frontend
    does not blur the comparison): predictions, loss, the gradient of EVERY parameter and the BatchNorm running statistics
    against (1) the oracle's autograd in fp64 on the CPU and (2) what the live reference produced in fp32 (train_net.npz).

    Gradient tolerance: with 2 clips the batch statistics of the coarse layers come from 16-64 rows, and ReLU / max-pool
    decisions that sit within rounding of a tie flip between any two fp32 implementations; the reference's OWN fp32
    gradients differ from the fp64 ones by 6e-5 (neck, layer4) rising to 7e-3 (layer1, the stem) on this fixture, and so do
    ours (tools/diag_train_grads.py prints both columns).  Hence: relative L2 error < 3e-2 for every parameter (a wrong
    kernel gives O(1)), < 5e-4 for everything that back-propagates before the first such flip (neck + layer4)."""
    import train_helpers as TH
    from yad_b200.train_engine import run_train_forward
    g = gold("train_net")
    x, tg = TH.train_inputs()
    preds_o, loss_o, grads_o, sd_o, xs = TH.oracle_train_step(ref_state_dict, x, tg, dtype=torch.float64)
    m = _train_model(ref_state_dict, cuda_dev)
    L_res = -(-320 * x.shape[-1] // 441)
    with torch.enable_grad():
        preds = run_train_forward(m, m._train_engine(), xs.to(cuda_dev).contiguous(), xs.shape[-1], L_res)
        loss, _ = _loss_fn()(preds, tg.to(cuda_dev))
        loss.backward()
    torch.cuda.synchronize()
    for i in range(3):
        np.testing.assert_allclose(preds[i].detach().cpu().numpy(), preds_o[i].numpy(), atol=1e-3, rtol=1e-4)
        np.testing.assert_allclose(preds[i].detach().cpu().numpy(), g[f"pred{i}"], atol=1e-3, rtol=1e-4)
    np.testing.assert_allclose(float(loss), float(loss_o), rtol=2e-5)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=2e-5)
    params = dict(m.named_parameters())
    assert sorted(params) == sorted(grads_o)
    ref_stats = {str(n): s for n, s in zip(g["grad_names"], g["grad_stats"])}
    worst = ("", 0.0)
    for k, p in params.items():
        assert p.grad is not None, k
        go = grads_o[k]
        scale = float(go.norm())
        if scale < 1e-9:      # conv biases in front of a batch-statistics BatchNorm (exactly zero gradient), unmatched anchors
            assert float(p.grad.double().norm()) < 1e-4, k
            continue
        r = _rel_l2(p.grad.cpu(), go)
        worst = max(worst, (k, r), key=lambda t: t[1])
        tight = k.startswith("multiscale_module.") or ".layer4." in k or k.endswith("_anchors")
        assert r < (5e-4 if tight else 3e-2), (k, r)
        np.testing.assert_allclose(TH.grad_stats(p.grad)[1], ref_stats[k][1], rtol=1e-3 if tight else 3e-2, err_msg=k)   # live reference
    for k in [k for k in g if k.startswith("grad:")]:
        tight = k[5:].startswith("multiscale_module.") or k.endswith("_anchors")
        if np.abs(g[k]).max() < 1e-6:       # zero-gradient parameter: the reference holds rounding noise, checked above
            continue
        np.testing.assert_allclose(params[k[5:]].grad.cpu().numpy(), g[k], rtol=5e-3,
                                   atol=(1e-3 if tight else 5e-2) * np.abs(g[k]).max(), err_msg=k)
    sd = m.state_dict()
    for k in [k[3:] for k in g if k.startswith("rm:")]:
        np.testing.assert_allclose(sd[k + ".running_mean"].cpu().numpy(), g["rm:" + k], atol=2e-6)
        np.testing.assert_allclose(sd[k + ".running_var"].cpu().numpy(), g["rv:" + k], rtol=2e-5)
        assert int(sd[k + ".num_batches_tracked"]) == int(g["nbt:" + k]) == 1
    print("worst relative gradient error:", worst)


def test_train_bottleneck_backbone_vs_oracle(variant_state_dict, cuda_dev):
    """``resnet_config.block: Bottleneck`` in train() mode (torchvision resnet.py:143-163 via modules/_backbone.py:128-138):
    predictions, loss, every parameter gradient and the BatchNorm running statistics against the oracle's autograd in fp64
    (same tolerances as the BasicBlock network: < 3e-2 relative L2 everywhere, tight on the neck)."""
    import train_helpers as TH
    from yad_b200.train_engine import run_train_forward
    sd_v, cfg = variant_state_dict("bottleneck")
    x, tg = TH.train_inputs()
    preds_o, loss_o, grads_o, sd_o, xs = TH.oracle_train_step(sd_v, x, tg, dtype=torch.float64)
    cfg = dict(cfg, dropout=0.0)
    m = yad_b200.AudioDetectionNetwork(2, config=cfg, train_dtype="f32")
    m.load_state_dict(sd_v)
    m = m.to(cuda_dev).train()
    L_res = -(-320 * x.shape[-1] // 441)
    with torch.enable_grad():
        preds = run_train_forward(m, m._train_engine(), xs.to(cuda_dev).contiguous(), xs.shape[-1], L_res)
        loss, _ = _loss_fn()(preds, tg.to(cuda_dev))
        loss.backward()
    torch.cuda.synchronize()
    for i in range(3):
        np.testing.assert_allclose(preds[i].detach().cpu().numpy(), preds_o[i].numpy(), atol=2e-3, rtol=2e-4)
    np.testing.assert_allclose(float(loss), float(loss_o), rtol=1e-4)
    params = dict(m.named_parameters())
    assert sorted(params) == sorted(grads_o)
    worst = ("", 0.0)
    for k, p in params.items():
        assert p.grad is not None, k
        go = grads_o[k]
        if float(go.norm()) < 1e-9:
            assert float(p.grad.double().norm()) < 1e-4, k
            continue
        r = _rel_l2(p.grad.cpu(), go)
        worst = max(worst, (k, r), key=lambda t: t[1])
        assert r < (2e-3 if k.startswith("multiscale_module.") else 3e-2), (k, r)
    sd = m.state_dict()
    for k in [k[:-13] for k in sd_o if k.endswith(".running_mean")]:
        np.testing.assert_allclose(sd[k + ".running_mean"].cpu().numpy(), sd_o[k + ".running_mean"].float().numpy(), atol=5e-6, rtol=1e-4)
        np.testing.assert_allclose(sd[k + ".running_var"].cpu().numpy(), sd_o[k + ".running_var"].float().numpy(), rtol=1e-4, atol=1e-7)
    print("worst relative gradient error:", worst)


def test_train_forward_end_to_end_and_eval_after(gold, ref_state_dict, cuda_dev):
    """model.train(); model(x) from PCM through the GPU frontend: loss within 2e-3 of the reference's; no_grad works;
    eval() afterwards uses the updated running statistics (engine re-packs)."""
    import train_helpers as TH
    g = gold("train_net")
    x, tg = TH.train_inputs()
    m = _train_model(ref_state_dict, cuda_dev)
    xd = x.to(cuda_dev)
    with torch.no_grad():
        p0 = m(xd, combine_scales=True)
    assert p0.shape == (2, (32 + 16 + 8) * 3, 5) and not p0.requires_grad
    m.load_state_dict(ref_state_dict)
    with torch.enable_grad():
        preds = m(xd)
        loss, met = _loss_fn()(preds, tg.to(cuda_dev))
        loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-3 * float(g["loss"])
    assert set(met) >= {"aggregate_loss", "mean_ciou", "conf_loss", "class_loss", "f1"}
    rm_before = ref_state_dict["feature_extractor.bn1.running_mean"]
    assert not torch.allclose(m.feature_extractor.bn1.running_mean.cpu(), rm_before)
    m.eval()
    with torch.no_grad():
        pe = m(xd, combine_scales=True)
    assert torch.isfinite(pe).all()


@pytest.mark.parametrize("train_dtype", ["f32", "tf32"])
def test_train_steps_reduce_loss_with_dropout(train_dtype, ref_state_dict, cuda_dev):
    """Five fused Adam + EMA steps on one batch (dropout 0.4 as in config.yaml): the loss goes down, gradients live in the
    flat arena, the EMA shadow moves."""
    import train_helpers as TH
    x, tg = TH.train_inputs()
    m = _train_model(ref_state_dict, cuda_dev, dropout=0.4, train_dtype=train_dtype)
    opt = yad_b200.FusedAdamEMA(m.parameters(), lr=1e-3, weight_decay=0.002, ema_momentum=0.002, use_ema=True)
    loss_fn = _loss_fn()
    xd, tgd = x.to(cuda_dev), tg.to(cuda_dev)
    losses = []
    ema0 = opt.ema.clone()
    for _ in range(5):
        with torch.enable_grad():
            loss, _ = loss_fn(m(xd), tgd)
            loss.backward()
        assert float(opt.grad.abs().sum()) > 0
        opt.step()
        opt.zero_grad()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
    assert not torch.equal(ema0, opt.ema)


# ------------------------------------------------------------------ TF32 tensor-core convolutions (csrc/conv_tf32.cu)
TF32_CASES = [
    # B, H, W, Cin, Cout, (kh, kw), (sh, sw), (ph, pw)
    (2, 8, 24, 64, 64, (3, 3), (1, 1), (1, 1)),
    (2, 8, 24, 64, 128, (3, 3), (2, 2), (1, 1)),
    (2, 8, 24, 64, 128, (1, 1), (2, 2), (0, 0)),
    (2, 16, 48, 64, 64, (7, 7), (2, 2), (3, 3)),
    (3, 1, 24, 15, 128, (3, 3), (1, 2), (1, 1)),
    (3, 1, 12, 128, 15, (3, 3), (1, 1), (1, 1)),
    (3, 1, 30, 15, 15, (1, 1), (1, 1), (0, 0)),
    (4, 1, 30, 512, 64, (1, 1), (1, 1), (0, 0)),
    (3, 2, 60, 256, 256, (3, 3), (1, 1), (1, 1)),
    (2, 7, 37, 96, 160, (3, 3), (1, 1), (1, 1)),
    (5, 4, 21, 128, 256, (3, 3), (2, 2), (1, 1)),
]


@pytest.mark.parametrize("case", TF32_CASES)
def test_conv_tf32_forward_dgrad_wgrad(case, ref_state_dict, cuda_dev):
    """tcgen05 kind::tf32 forward / data gradient / weight gradient (through TrainEngine.conv, i.e. with the host-side tap
    lists and weight packing) against fp32 PyTorch autograd.  TF32 keeps 10 mantissa bits per operand: tolerance 2e-3 of
    the tensor's scale (cuDNN's default TF32 convolutions differ from fp32 by the same amount)."""
    from yad_b200.train_engine import _T
    B, H, W, Cin, Cout, k, s, p = case
    g = torch.Generator().manual_seed(abs(hash(case)) % 9973)
    conv = torch.nn.Conv2d(Cin, Cout, k, s, p, bias=True).to(cuda_dev)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (Cin * k[0] * k[1]) ** 0.5)
        conv.bias.copy_(torch.randn(Cout, generator=g))
    x = torch.randn(B, Cin, H, W, generator=g).to(cuda_dev).requires_grad_(True)
    y = F.conv2d(x, conv.weight, conv.bias, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g).to(cuda_dev)
    gw, gb, gx = torch.autograd.grad(y, (conv.weight, conv.bias, x), dy)
    m = _train_model(ref_state_dict, cuda_dev, train_dtype="tf32")
    eng = m._train_engine()
    eng._tape, eng._grads, eng._bn_counters = [], {}, []
    xt = eng._new(B, H, W, Cin)
    xt.buf[..., :Cin] = x.detach().permute(0, 2, 3, 1)
    out = eng.conv(xt, conv)
    assert eng._conv_tf32.__name__ and len(eng._plans) == 1      # the tensor-core path ran
    torch.cuda.synchronize()
    tol = lambda ref: 2e-3 * float(ref.abs().max())          # noqa: E731
    np.testing.assert_allclose(out.buf[..., :Cout].permute(0, 3, 1, 2).cpu().numpy(), y.detach().cpu().numpy(), atol=tol(y))
    if out.ld > Cout:
        assert float(out.buf[..., Cout:].abs().max()) == 0.0       # pad channels stay zero
    prev = torch.randn(B, H, W, Cin, generator=g).to(cuda_dev)     # an existing gradient to accumulate onto
    eng._grad(out).buf[..., :Cout] = dy.permute(0, 2, 3, 1)
    eng._grad(xt).buf[..., :Cin] = prev
    conv.weight.grad = torch.ones_like(conv.weight)
    conv.bias.grad = None
    for fn in reversed(eng._tape):
        fn()
    torch.cuda.synchronize()
    dx = eng._grad(xt).buf
    np.testing.assert_allclose((dx[..., :Cin] - prev).permute(0, 3, 1, 2).cpu().numpy(), gx.cpu().numpy(), atol=tol(gx))
    if xt.ld > Cin:
        assert float(dx[..., Cin:].abs().max()) == 0.0
    np.testing.assert_allclose((conv.weight.grad - 1).cpu().numpy(), gw.cpu().numpy(), atol=tol(gw))
    np.testing.assert_allclose(conv.bias.grad.cpu().numpy(), gb.cpu().numpy(), rtol=1e-4, atol=1e-4)
    eng._tape, eng._grads = None, {}


def test_train_network_tf32_vs_cudnn_tf32_yardstick(gold, ref_state_dict, cuda_dev):
    """The default train mode (tcgen05 kind::tf32 convolutions).  With 2 clips this fixture is hypersensitive to operand
    rounding (batch statistics over 16-64 rows): PyTorch's OWN TF32 convolutions (cuDNN, torch's default for fp32 training)
    move the gradients of the oracle by 25 % (median, relative L2 vs fp64) on it.  So the yardstick is measured, not assumed:
    the oracle runs on the GPU with allow_tf32 and our error against the fp64 oracle must not exceed 1.5x its error (median
    and maximum over the parameters); the loss must agree with fp64 to 2e-3."""
    import train_helpers as TH
    from yad_b200.train_engine import run_train_forward
    x, tg = TH.train_inputs()
    preds_o, loss_o, grads_o, _, xs = TH.oracle_train_step(ref_state_dict, x, tg, dtype=torch.float64)
    # yardstick: oracle CNN on the GPU with cuDNN TF32, decode + loss on the CPU
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        sdc = {k: v.to(cuda_dev) for k, v in ref_state_dict.items()}
        names = TH.param_names(sdc)
        with torch.enable_grad():
            for k in names:
                sdc[k].requires_grad_(True)
            O._BN_TRAINING[0] = True
            try:
                heads = O.neck(sdc, O.backbone(sdc, xs.to(cuda_dev), TH.TRAIN_CFG["block_layers"]))
            finally:
                O._BN_TRAINING[0] = False
            sd_cpu = {k: v.cpu() for k, v in sdc.items()}       # differentiable copies: anchors keep their graph
            L_res = -(-320 * x.shape[-1] // 441)
            pr = O.decode([h.cpu() for h in heads], sd_cpu, L_res, xs.shape[-1], 2, TH.TRAIN_CFG, combine_scales=False)
            loss_y, _ = O.detection_loss(pr, tg, O.DEFAULT_CONFIG["anchors"], 2)
            loss_y.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    live = [k for k in names if float(grads_o[k].norm()) > 1e-9 and not k.endswith("_anchors")]
    err_y = sorted(_rel_l2(sdc[k].grad.cpu(), grads_o[k]) for k in live)
    m = _train_model(ref_state_dict, cuda_dev, train_dtype="tf32")
    with torch.enable_grad():
        preds = run_train_forward(m, m._train_engine(), xs.to(cuda_dev).contiguous(), xs.shape[-1], L_res)
        loss, _ = _loss_fn()(preds, tg.to(cuda_dev))
        loss.backward()
    assert abs(float(loss) - float(loss_o)) < 2e-3 * float(loss_o), (float(loss), float(loss_o))
    params = dict(m.named_parameters())
    err = sorted(_rel_l2(params[k].grad.cpu(), grads_o[k]) for k in live)
    med, med_y = err[len(err) // 2], err_y[len(err_y) // 2]
    print(f"tf32 gradient error vs fp64: ours median {med:.3f} max {err[-1]:.3f}; cuDNN TF32 median {med_y:.3f} max {err_y[-1]:.3f}")
    assert med <= 1.5 * med_y + 1e-3 and err[-1] <= 1.5 * err_y[-1] + 1e-3
    for k in names:
        if float(grads_o[k].norm()) <= 1e-9:
            assert float(params[k].grad.double().norm()) < 1e-3, k


def test_train_step_cuda_graphs_equal_eager(ref_state_dict, cuda_dev):
    """model.train_graphs = True: the third call of a shape captures frontend + forward and the backward into CUDA graphs,
    later calls replay them.  Same predictions / loss / gradients as the eager path (dropout 0; fp32 red.add order differs)."""
    import train_helpers as TH
    x, tg = TH.train_inputs()
    xd, tgd = x.to(cuda_dev), tg.to(cuda_dev)
    out = {}
    for mode in ("eager", "graph"):
        m = _train_model(ref_state_dict, cuda_dev, train_dtype="tf32")
        opt = yad_b200.FusedAdamEMA(m.parameters(), lr=1e-3, weight_decay=0.002)
        m.train_graphs = mode == "graph"
        loss_fn = _loss_fn()
        rec = []
        for it in range(5):
            m.load_state_dict(ref_state_dict)          # same parameters AND running statistics at every iteration
            with torch.enable_grad():
                preds = m(xd)
                loss, _ = loss_fn(preds, tgd)
                loss.backward()
            rec.append((float(loss), opt.grad.clone(), m.feature_extractor.bn1.running_mean.clone(),
                        int(m.feature_extractor.bn1.num_batches_tracked)))
            opt.zero_grad()
        out[mode] = rec
        if mode == "graph":
            assert any(e["g"] is not None for e in m._train_engine()._graphs.values()), "the graphs were not captured"
    print("losses eager", [r[0] for r in out["eager"]], "graph", [r[0] for r in out["graph"]])
    print("grad rel: eager3-vs-eager4", _rel_l2(out["eager"][3][1].cpu(), out["eager"][4][1].cpu()), "graph3-vs-eager3",
          _rel_l2(out["graph"][3][1].cpu(), out["eager"][3][1].cpu()))
    for it in (3, 4):        # replayed iterations
        le, ge, re_, _ = out["eager"][it]
        lg, gg, rg, nb = out["graph"][it]
        assert abs(le - lg) < 1e-6 * abs(le), (le, lg)     # the forward is bit-reproducible (no split-K atomics in it)
        # backward: fp32 red.add order (wgrad, split-K dgrad) varies; TF32 operand truncation turns that 1e-7 noise into ~1e-3
        assert _rel_l2(gg.cpu(), ge.cpu()) < 1e-2
        np.testing.assert_allclose(rg.cpu().numpy(), re_.cpu().numpy(), atol=1e-5)
    assert out["graph"][4][3] == 1                          # load_state_dict reset the counter, the replay incremented it


def test_collate_batch_vs_live_reference(gold, cuda_dev):
    """GPU batch builder == the live reference's __getitem__ + collate_fn (fixture collate.npz), bit exact;
    the int16 entry equals the oracle fed the de-quantised waveforms."""
    from test_oracle_golden import _collate_inputs
    g = gold("collate")
    clips, segs, gmins, sr, dur = _collate_inputs()
    tg = [yad_b200.clip_targets(s, s[0][0], s[-1][1], c.shape[-1], sr, dur, -100, gm) for c, s, gm in zip(clips, segs, gmins)]
    audio, targets = yad_b200.collate_batch(clips, tg, sr, dur, cuda_dev)
    np.testing.assert_array_equal(audio.cpu().numpy(), g["audio"])
    np.testing.assert_array_equal(targets.cpu().numpy(), g["targets"])
    ci = [(c * 32767).round().to(torch.int16) for c in clips]
    a16, t16 = yad_b200.collate_batch(ci, tg, sr, dur, cuda_dev)
    ref, _ = O.getitem_collate([c.float() / 32768.0 for c in ci], segs, sr, dur, -100, gmins)
    np.testing.assert_allclose(a16.cpu().numpy(), ref.numpy(), atol=1e-7, rtol=0)      # 3-channel mean: one rounding apart at most
    np.testing.assert_array_equal(t16.cpu().numpy(), g["targets"])
    with pytest.raises(ValueError):
        yad_b200.collate_batch([torch.zeros(1, sr * dur + 1)], [tg[0]], sr, dur, cuda_dev)
