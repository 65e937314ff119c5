"""Training-side pieces of the hot path on sm_100a kernels."""
from __future__ import annotations

import ctypes as C
import math
from copy import deepcopy
from typing import List, Sequence, Tuple, Union

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def build_target_by_scale(targets: torch.Tensor, fmap_shape: Union[int, torch.Size], anchors: Union[Sequence[float], torch.Tensor],
                          anchor_threshold: float = 4.0, sample_duration: float = 60, edge_threshold: float = 0.5,
                          ) -> Tuple[List[torch.Tensor], torch.Tensor, torch.Tensor]:
    """Drop-in for ``AudioDataset.build_target_by_scale`` (reference: dataset.py:286-365).

    targets [T,4] f32 = (batch_idx, cls, centre_s, dur_s) on a CUDA device.  Returns
    ([batch_idx, grid_idx, anchor_idx] i64, classes i64, cw [M,2] f32) in the reference's row order."""
    if not targets.is_cuda:
        raise RuntimeError("yad_b200.build_target_by_scale needs CUDA tensors (no CPU fallback)")
    if isinstance(fmap_shape, torch.Size):
        fmap_shape = fmap_shape[0]
    dev = targets.device
    lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
    anc = torch.as_tensor(anchors, dtype=torch.float32).to(dev).contiguous()
    t = targets.contiguous().float()
    T, A = t.shape[0], anc.shape[0]
    if T == 0:
        z = torch.zeros(0, device=dev, dtype=torch.int64)
        return [z, z.clone(), z.clone()], z.clone(), torch.zeros((0, 2), device=dev, dtype=torch.float32)
    cap = max(1, 3 * A * T)
    bi = torch.empty(cap, device=dev, dtype=torch.int64)
    gi = torch.empty(cap, device=dev, dtype=torch.int64)
    ai = torch.empty(cap, device=dev, dtype=torch.int64)
    cl = torch.empty(cap, device=dev, dtype=torch.int64)
    cw = torch.empty((cap, 2), device=dev, dtype=torch.float32)
    n = torch.zeros(1, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        rc = lib.yad_build_targets(t.data_ptr(), T, anc.data_ptr(), A, int(fmap_shape), float(anchor_threshold),
                                   float(sample_duration), float(edge_threshold), bi.data_ptr(), gi.data_ptr(), ai.data_ptr(),
                                   cl.data_ptr(), cw.data_ptr(), n.data_ptr(), _stream(dev))
    _lib.check(rc, "build_targets")
    M = int(n.item())     # data-dependent shape, as in the reference (boolean-mask indexing syncs too)
    return [bi[:M], gi[:M], ai[:M]], cl[:M], cw[:M]


def clip_targets(segments, audio_start: float, audio_end: float, n_samples: int, sample_rate: int, sample_duration: float,
                 ignore_index: int = -100, gmin: float = 0.0) -> torch.Tensor:
    """Labels of ONE clip as ``AudioDataset.__getitem__`` builds them (reference: dataset.py:123-160): rows
    (0, class, centre_s, dur_s) from (start_s, end_s, class_idx) segments, shifted by the group minimum, plus the
    ignore-label row that covers the zero padding of a clip shorter than ``sample_duration``."""
    import numpy as np
    seg = np.asarray(segments, dtype=np.float64).reshape(-1, 3)
    times = seg[:, :2].copy() - gmin
    a_start, a_end = audio_start - gmin, audio_end - gmin
    times[:, 1] = times[:, 1] - times[:, 0]
    times[:, 0] = times[:, 0] + times[:, 1] / 2
    labels = torch.cat((torch.from_numpy(seg[:, 2].astype(np.int64))[:, None], torch.from_numpy(times).to(torch.float32)), dim=-1)
    if n_samples < sample_duration * sample_rate:
        pad_dur = (a_start + sample_duration) - a_end
        pad_c = a_end + pad_dur / 2
        labels = torch.cat((labels, torch.tensor([[float(ignore_index), pad_c, pad_dur]], dtype=labels.dtype)), dim=0)
    out = torch.zeros((labels.shape[0], 4), dtype=labels.dtype)
    out[:, 1:] = labels
    return out


def collate_batch(clips, targets, sample_rate: int, sample_duration: float, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """GPU-side ``AudioDataset.collate_fn`` (reference: dataset.py:132-155,276-283) for clips that are still ragged: ``clips`` is
    a list of host tensors [C_i, n_i] (fp32, or int16 PCM straight from the decoder), ``targets`` the matching list of
    ``clip_targets`` results.  ONE packed host->device copy, one kernel (channel mean, zero padding to
    ``sample_duration * sample_rate``), targets concatenated with their batch index.  Returns (audio [B,1,L] f32, targets [T,4] f32)
    on ``device``."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("yad_b200.collate_batch builds the batch on a CUDA device (no CPU fallback)")
    lib = _lib.init(device.index if device.index is not None else torch.cuda.current_device())
    B = len(clips)
    L = int(sample_duration * sample_rate)
    dt = clips[0].dtype
    if dt not in (torch.float32, torch.int16) or any(c.dtype != dt for c in clips):
        raise ValueError("collate_batch: clips must all be float32 or all be int16")
    shapes = [(1, c.shape[0]) if c.ndim == 1 else tuple(c.shape) for c in clips]
    for ch, n in shapes:
        if n > L:
            raise ValueError(f"audio sample is more than {sample_duration}, ensure that the specified sample rate value "
                             f"({sample_rate}) is correct")        # dataset.py:133-137
    sizes = [ch * n for ch, n in shapes]
    offs = [0]
    for sz in sizes[:-1]:
        offs.append(offs[-1] + sz)
    packed = torch.empty(sum(sizes), dtype=dt).pin_memory()
    for c, o, sz in zip(clips, offs, sizes):
        packed[o:o + sz] = c.reshape(-1)
    meta = torch.tensor(offs, dtype=torch.int64), torch.tensor([n for _, n in shapes], dtype=torch.int32), \
        torch.tensor([ch for ch, _ in shapes], dtype=torch.int32)
    pd = packed.to(device, non_blocking=True)
    od, nd, cd = (m.to(device, non_blocking=True) for m in meta)
    audio = torch.empty((B, 1, L), device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        rc = lib.yad_collate_clips(pd.data_ptr(), 1 if dt == torch.int16 else 0, od.data_ptr(), nd.data_ptr(), cd.data_ptr(), B, L,
                                   audio.data_ptr(), _stream(device))
    _lib.check(rc, "collate_clips")
    tg = []
    for i, t in enumerate(targets):
        t = t.clone()
        t[:, 0] = i
        tg.append(t)
    return audio, torch.cat(tg, dim=0).to(device, non_blocking=True)


class _DetectionLossFn(torch.autograd.Function):
    """loss = f(preds); the kernels return d loss / d preds, so backward is a scale by the upstream gradient."""

    @staticmethod
    def forward(ctx, mod, targets, *preds):
        loss, grads, metrics = mod._run(preds, targets)
        ctx.save_for_backward(*grads)
        mod._last_metrics = metrics
        return loss

    @staticmethod
    def backward(ctx, g):
        return (None, None) + tuple(g * t for t in ctx.saved_tensors)


class AudioDetectionLoss(torch.nn.Module):
    """Drop-in for the reference ``AudioDetectionLoss`` (modules/_loss.py:39-191) on sm_100a kernels: same constructor,
    ``forward((sm, md, lg), targets) -> (loss, metrics dict)`` with the same 10 metric keys; ``loss`` is differentiable
    with respect to the three prediction tensors (CUDA, fp32).  All branches of the reference constructor (:74-81):
    ``multi_label`` BCE class loss with label smoothing (the train_config default) or ``CrossEntropyLoss(weight=class_weights)``,
    BCEWithLogits objectness or ``FocalLoss(alpha, gamma)`` (yad_loss_scale_ex)."""

    SCALE_W = (4.0, 2.0, 1.0)

    def __init__(self, anchors_dict, num_classes, anchor_t=4.0, edge_t=0.5, sample_duration=60, box_w=1.0, conf_w=1.0,
                 class_w=1.0, multi_label=False, class_weights=None, label_smoothing=0, batch_scale_loss=False, alpha=None,
                 gamma=None, ignore_index=-100):
        super().__init__()
        self.multi_label = bool(multi_label)
        self.focal = (float(alpha), float(gamma)) if (alpha and gamma) else (0.0, 0.0)      # same truthiness test as the reference
        self.class_weights = None if (class_weights is None or multi_label) else torch.as_tensor(class_weights, dtype=torch.float32)
        self.anchors_dict, self.num_classes = anchors_dict, int(num_classes)
        self.anchor_t, self.edge_t, self.sample_duration = anchor_t, edge_t, sample_duration
        self.box_w, self.conf_w, self.class_w = float(box_w), float(conf_w), float(class_w)
        self.label_smoothing, self.batch_scale_loss, self.ignore_index = float(label_smoothing), bool(batch_scale_loss), int(ignore_index)
        self._last_metrics = {}

    def _run(self, preds, targets):
        dev = preds[0].device
        if dev.type != "cuda":
            raise RuntimeError("yad_b200.AudioDetectionLoss needs CUDA tensors (no CPU fallback)")
        lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
        nc = self.num_classes
        bscale = float(preds[-1].shape[0]) if self.batch_scale_loss else 1.0
        accs = torch.zeros((3, 16), device=dev, dtype=torch.float64)
        cwt = None if self.class_weights is None else self.class_weights.to(dev).contiguous()
        if cwt is not None and cwt.numel() != nc:
            raise ValueError(f"class_weights has {cwt.numel()} entries, expected {nc}")
        confs = torch.zeros((3, nc, nc), device=dev, dtype=torch.int32)
        grads, Ms, Ns = [], [], []
        with torch.cuda.device(dev):
            for s, (p, name) in enumerate(zip(preds, ("sm", "md", "lg"))):
                p = p.detach().contiguous().float()
                B, G, A, E = p.shape
                if E != 3 + nc:
                    raise ValueError(f"prediction tensor has {E} channels, expected {3 + nc}")
                (bi, gi, ai), cl, cw = build_target_by_scale(targets, G, self.anchors_dict[name], self.anchor_t,
                                                             self.sample_duration, self.edge_t)
                M = int(bi.shape[0])
                owner = torch.empty(B * G * A, device=dev, dtype=torch.int32)
                ciou = torch.empty(max(M, 1), device=dev, dtype=torch.float32)
                g = torch.empty_like(p)
                rc = lib.yad_loss_scale_ex(p.data_ptr(), B, G, A, nc, _lib.ptr(bi), _lib.ptr(gi), _lib.ptr(ai), _lib.ptr(cl), _lib.ptr(cw),
                                           M, self.box_w * bscale, self.conf_w * self.SCALE_W[s] * bscale, self.class_w * bscale,
                                           self.label_smoothing, self.ignore_index, 0 if self.multi_label else 1, _lib.ptr(cwt),
                                           self.focal[0], self.focal[1], owner.data_ptr(), ciou.data_ptr(),
                                           confs[s].data_ptr(), accs[s].data_ptr(), g.data_ptr(), _stream(dev))
                _lib.check(rc, f"loss_scale {name}")
                grads.append(g)
                Ms.append(M)
                Ns.append(B * G * A)
        a = accs.cpu().numpy()            # one 384-byte read; the reference syncs on every .item() too
        cm = confs.cpu().numpy()
        nan = float("nan")
        lbox = lconf = lcls = 0.0
        per = {k: [] for k in ("mean_ciou", "conf_loss", "avg_pos_conf", "avg_neg_conf", "class_loss", "accuracy", "f1", "precision", "recall")}
        for s in range(3):
            M, N, nvalid = Ms[s], Ns[s], a[s, 7]
            box = a[s, 0] / M if M else nan
            conf = a[s, 2] / N
            if self.multi_label:
                cls = a[s, 3] / (nvalid * nc) if nvalid else nan
            else:            # CrossEntropyLoss: weighted mean; no class-valid match (or zero weight sum) -> nan, like torch
                cls = a[s, 3] / a[s, 8] if (nvalid and a[s, 8]) else nan
            lbox += 0.0 if box != box else box          # handle_nan (modules/_loss.py:178)
            lconf += self.SCALE_W[s] * conf
            lcls += 0.0 if cls != cls else cls
            per["mean_ciou"].append(a[s, 1] / M if M else nan)
            per["conf_loss"].append(conf)
            per["avg_pos_conf"].append(a[s, 4] / M if M else nan)
            per["avg_neg_conf"].append(a[s, 5] / a[s, 6] if a[s, 6] else nan)
            per["class_loss"].append(cls)
            acc_, f1_, pr_, rc_ = _macro_metrics(cm[s]) if nvalid else (nan, nan, nan, nan)
            per["accuracy"].append(acc_)
            per["f1"].append(f1_)
            per["precision"].append(pr_)
            per["recall"].append(rc_)
        # the reference averages the three scales with pandas DataFrame.mean(), which SKIPS NaN (modules/_loss.py:101-110): a
        # scale without matches does not poison the metric; NaN only when all three are NaN
        met = {}
        for k, vals in per.items():
            ok = [float(v) for v in vals if v == v]
            met[k] = sum(ok) / len(ok) if ok else nan
        loss_v = (self.box_w * lbox + self.conf_w * lconf + self.class_w * lcls) * bscale
        met = {"aggregate_loss": float(loss_v), **{k: float(v) for k, v in met.items()}}
        return torch.tensor(loss_v, device=dev, dtype=torch.float32), grads, met

    def forward(self, preds, targets):
        loss = _DetectionLossFn.apply(self, targets, *preds)
        return loss, dict(self._last_metrics)


def _macro_metrics(cm):
    """accuracy and sklearn's macro precision / recall / f1 (labels = those present in targets or predictions,
    zero_division -> 0) from a confusion matrix [target, predicted]."""
    import numpy as np
    cm = np.asarray(cm, dtype=np.float64)
    tot = cm.sum()
    tp = np.diag(cm)
    sup_t, sup_p = cm.sum(1), cm.sum(0)
    present = (sup_t + sup_p) > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        prec = np.where(sup_p > 0, tp / sup_p, 0.0)
        rec = np.where(sup_t > 0, tp / sup_t, 0.0)
        f1 = np.where(prec + rec > 0, 2 * prec * rec / (prec + rec), 0.0)
    return tp.sum() / tot, f1[present].mean(), prec[present].mean(), rec[present].mean()


class FusedAdamEMA:
    """torch.optim.Adam (L2 weight decay; train.py:83-90, config.yaml:75-80) fused with the EMA shadow update
    (smoothener/_ema.py:20-26) over ONE flat fp32 arena: 1 launch per step instead of ~10 per tensor.

    ``params`` are re-pointed into the arena (views), so the model keeps working unchanged."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, ema_momentum: float = 0.0,
                 ema_N: int = 2000, use_ema: bool = False):
        self.params = [p for p in params]
        # like torch.optim.Adam (train.py:83-90), parameters that never receive a gradient are left alone: only tensors with
        # requires_grad live in the arena (a frozen tensor - e.g. the anchors when train_anchors is false - would otherwise be
        # stepped with g = weight_decay * p and drift).  state_dict() still indexes ALL parameters, as Adam's does.
        self.active = [i for i, p in enumerate(self.params) if p.requires_grad]
        if not self.active:
            raise ValueError("FusedAdamEMA: no parameter requires a gradient")
        dev = self.params[self.active[0]].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamEMA needs CUDA parameters (no CPU fallback)")
        self.dev = dev
        self.lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.step_count = 0
        n = sum(self.params[i].numel() for i in self.active)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.offsets = {}
        o = 0
        for i in self.active:
            p = self.params[i]
            k = p.numel()
            self.flat[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + k].view_as(p.data)
            p.grad = self.grad[o:o + k].view_as(p.data)
            self.offsets[i] = (o, k)
            o += k
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.use_ema = use_ema
        self.ema = self.flat.clone() if use_ema else None
        self.ema_momentum, self.ema_N = ema_momentum, ema_N

    def ema_m(self, n: int) -> float:
        return 1 - ((1 - self.ema_momentum) * (1 - math.exp(-n / self.ema_N)))

    def zero_grad(self):
        self.grad.zero_()

    def allreduce_grads(self):
        """Average the flat gradient arena across the data-parallel ranks (one bucketed NCCL all-reduce).
        Serial form: call it after ``loss.backward()``.  See ``overlap_allreduce`` for the overlapped form."""
        from .parallel import allreduce_mean_
        allreduce_mean_(self.grad)

    def overlap_allreduce(self, model) -> None:
        """Bucket the gradient all-reduce behind the backward (pipeline/_trainer.py:94-108 under data parallelism):
        the train engine reports each gradient bucket as soon as its last kernel is enqueued (neck + layer4 = 75 % of the bytes
        after ~15 % of the backward) and this optimizer launches that bucket's asynchronous all-reduce on NCCL's stream; call
        ``wait_allreduce()`` between ``loss.backward()`` and ``step()``.  The buckets are contiguous ranges of the flat arena."""
        eng = model._train_engine()
        pos = {id(self.params[i]): oc for i, oc in self.offsets.items()}
        self._buckets = []
        for plist in eng.bucket_params():
            rng = [pos[id(p)] for p in plist if id(p) in pos]
            if not rng:
                self._buckets.append(None)
                continue
            # contiguous runs of the bucket's parameters in the arena (module order puts the stem's conv2 after layer4, so two of
            # the three buckets are two runs each)
            runs = []
            for o, k in sorted(rng):
                if runs and runs[-1][1] == o:
                    runs[-1][1] = o + k
                else:
                    runs.append([o, o + k])
            self._buckets.append([tuple(r) for r in runs])
        from .parallel import BucketedAllReduce
        self._bar = BucketedAllReduce(self.grad, self._buckets)
        eng.on_bucket = self._bar.ready

    def wait_allreduce(self) -> None:
        """Make the current stream wait for the bucket all-reduces launched during the backward."""
        bar = getattr(self, "_bar", None)
        if bar is not None:
            bar.wait()

    def step(self):
        self.step_count += 1
        m = self.ema_m(self.step_count) if self.use_ema else 0.0
        with torch.cuda.device(self.dev):
            rc = self.lib.yad_adam_ema_step(self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                            _lib.ptr(self.ema), self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                            self.wd, self.step_count, m, _stream(self.dev))
        _lib.check(rc, "adam_ema_step")
        _lib.param_epoch += 1        # the packed-weight caches of the engines key on this


    # ---- checkpoint plumbing (pipeline/_trainer.py:38-53 stores optimizer.state_dict() under "optimizer_params")
    def state_dict(self) -> dict:
        """The layout of ``torch.optim.Adam.state_dict()`` (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``, one param group),
        so checkpoints written here load into the reference's Adam and vice versa."""
        state = {}
        for i, (o, k) in self.offsets.items():
            p = self.params[i]
            if self.step_count > 0:
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.m[o:o + k].view_as(p.data).clone(),
                            "exp_avg_sq": self.v[o:o + k].view_as(p.data).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("FusedAdamEMA.load_state_dict: expected one param group covering every parameter")
        g = groups[0]
        if g.get("amsgrad") or g.get("maximize") or g.get("decoupled_weight_decay"):
            raise NotImplementedError("FusedAdamEMA: amsgrad / maximize / decoupled weight decay are not built")
        self.lr, self.betas, self.eps, self.wd = float(g["lr"]), tuple(g["betas"]), float(g["eps"]), float(g["weight_decay"])
        steps = set()
        self.m.zero_()
        self.v.zero_()
        for i, (o, k) in self.offsets.items():
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is not None:
                self.m[o:o + k].copy_(st["exp_avg"].reshape(-1))
                self.v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise NotImplementedError("FusedAdamEMA: per-parameter step counts differ (one fused bias correction per step)")
        self.step_count = steps.pop() if steps else 0


def save_checkpoint(path: str, model, optimizer, ema_smoothener=None) -> None:
    """TrainerPipeline.save_model (pipeline/_trainer.py:38-47): {"network_params", "optimizer_params"}; the EMA weights replace
    the live ones when a smoothener is in use."""
    import os
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    net = model.state_dict() if not ema_smoothener else ema_smoothener.get_ema_state_dict()
    torch.save({"network_params": net, "optimizer_params": optimizer.state_dict()}, path)


def load_checkpoint(path: str, model, device=None, optimizer=None):
    """TrainerPipeline.load_model (pipeline/_trainer.py:49-53) / inference.py:29-31; optionally resumes the optimizer too."""
    import os
    if not os.path.exists(path):
        raise OSError(f"model is yet to be saved in path: {path}")
    saved = torch.load(path, map_location=device)
    out = model.load_state_dict(saved["network_params"])
    _lib.param_epoch += 1
    if optimizer is not None and "optimizer_params" in saved:
        optimizer.load_state_dict(saved["optimizer_params"])
    return out


class EMAParamsSmoothener:
    """Drop-in for smoothener/_ema.py:7-32 (parameters only; warm-up momentum).  The update is two ``torch._foreach`` calls over
    the parameter lists (ATen multi-tensor kernels, not a kernel of this library); the train step's own EMA is fused into
    ``FusedAdamEMA`` (``yad_adam_ema_step``).  Every update bumps ``_lib.param_epoch`` so an ``ema_model`` in eval mode re-packs
    its weights instead of running stale ones (in-place ``.data`` writes do not change ``tensor._version``)."""

    def __init__(self, model, momentum: float = 0.002, num_updates: int = 0, N: int = 2_000):
        self.model = model
        self.ema_model = deepcopy(model)
        self.ema_model.eval()
        self.num_updates = num_updates
        self.momentum_ = lambda n: 1 - ((1 - momentum) * (1 - math.exp(-n / N)))
        for p in self.ema_model.parameters():
            p.requires_grad_(False)

    def update(self):
        self.num_updates += 1
        m = self.momentum_(self.num_updates)
        with torch.no_grad():
            ema = [e.data for e in self.ema_model.parameters() if e.dtype.is_floating_point]
            src = [p.data for e, p in zip(self.ema_model.parameters(), self.model.parameters()) if e.dtype.is_floating_point]
            torch._foreach_mul_(ema, 1 - m)
            torch._foreach_add_(ema, src, alpha=m)
        _lib.param_epoch += 1

    def get_ema_state_dict(self):
        return self.ema_model.state_dict()

    def load_state_dict(self, state_dict):
        self.ema_model.load_state_dict(state_dict)
        _lib.param_epoch += 1
