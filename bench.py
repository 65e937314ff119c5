#!/usr/bin/env python
"""bench.py - audio-seconds/sec of the detection hot path on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (oracle port) on the host cores

A "step" = one pass of the whole hot path over one batch of synthetic 60 s clips per GPU:
PCM -> fused log-mel/MFCC frontend -> stem + ResNet + RepBi-PAN neck (deploy form, bf16 tcgen05) -> anchor
decode -> per-clip segment NMS.  Prints ONE JSON line (rank 0).  Per-GPU work is fixed (weak scaling);
clips are sharded across ranks with no data-path collective.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

CLIP_SECONDS = 60.0
CLIP_SAMPLES = 1323000
FRONTEND_BYTES_PER_CLIP = 4 * CLIP_SAMPLES + 2 * 32 * 960 * 4          # PCM read + feature write (whole frontend)
MEL_KERNEL_BYTES_PER_CLIP = 4 * CLIP_SAMPLES + 32 * 960 * 4            # frontend_mel_kernel alone: PCM read + mel-power write


def load_mel_kernel_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of frontend_mel_kernel per clip, parsed from the newest committed
    `ncu --set full` summary of the 512-clip launch (profiles/rNN_ncu_full_frontend_mel_b512_vM.txt); (None, None) if absent."""
    import glob
    import re

    def key(path):
        m = re.search(r"r(\d+)_ncu_full_frontend_mel_b512_v(\d+)", path)
        return (int(m.group(1)), int(m.group(2))) if m else (-1, -1)
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_frontend_mel_b512_v*.txt")), key=key)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in reversed(files):
        tot, seen = 0.0, 0
        with open(path) as f:
            for line in f:
                m = re.match(r"dram__bytes_(read|write)\.sum \[(\w+)\]: ([0-9.eE+-]+)", line.strip())
                if m and seen < 2:
                    tot += float(m.group(3)) * unit.get(m.group(2), 1.0)
                    seen += 1
        if seen == 2:
            return tot / 512.0, os.path.relpath(path, ROOT)
    return None, None


CNN_FLOP_PER_CLIP = 2 * 1150923632                                     # useful MACs of the reference graph, deploy form
# executed by this implementation: conv1 (45.5 MMAC) and conv2 (342.8 MMAC) run as ONE composite 19x19 stride-4 convolution
# (8 x 240 pixels x 722 taps x 64 channels = 88.7 MMAC)
CNN_FLOP_EXECUTED_PER_CLIP = 2 * (1150923632 - 45500000 - 342800000 + 8 * 240 * 722 * 64)
METRIC = "audio-seconds/sec (mel+RepVGG fwd+decode/NMS)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock + throttle reasons under the benchmark's load (B200_PROFILING.md recipe), sampled from the MAIN thread while
    enqueued steps execute: the warm-up steps right before the timed region, an untimed repeat of the K timed steps right after
    it, and the end-to-end loop.  Queries INSIDE the timed region were measured to corrupt it, whichever way they are made:
    a background nvidia-smi every 0.2 s -> 3.99 or 4.6 ms per step depending on whether a query landed in the 40 ms region;
    NVML in a thread every 25 ms -> 6.4-9.2 ms per step; one NVML query from the main thread after the K steps were enqueued ->
    3.9 ms in most runs but 13.4 ms in two of twelve (a ~95 ms device stall).  Uses NVML in-process
    (nvidia_ml_py: the data source of `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`), falls back to
    the nvidia-smi subprocess."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples = index, []
        self.nvml = self.handle = None
        self._init_done = False

    def _init(self):
        # lazily, at the first sample: not even nvmlInit() runs before the timed regions are over
        self._init_done = True
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:  # noqa: BLE001
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = pynvml
        except Exception:  # noqa: BLE001
            self.nvml = None

    def sample(self):
        """One sample; call it while the device is busy with already-enqueued work."""
        if not self._init_done:
            self._init()
        try:
            if self.nvml is not None:
                n, h = self.nvml, self.handle
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                r = int(get(h))
                act = lambda bit: "Active" if r & bit else "Not Active"   # noqa: E731
                # nvml.h: HwSlowdown 0x8, HwThermalSlowdown 0x40, SwThermalSlowdown 0x20, SwPowerCap 0x4
                self.samples.append([str(sm), str(mx), act(0x8), act(0x40), act(0x20), act(0x4)])
            else:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([s.strip() for s in o.split(",")])
        except Exception:  # noqa: BLE001
            pass

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi",
                "when": "after every timed measurement: while the device executed untimed repeats of the K timed steps (and of one "
                        "end-to-end step); no monitoring query runs before or inside a timed region - NVML / nvidia-smi queries "
                        "stall the device (see ClockSampler)"}


def synth_clips_device(B: int, dev, seed: int):
    """[B,1,L] f32 synthetic clips generated on the device: 0.1*noise + 4 gated tone bursts per clip, every 8th
    clip with a digital-silence tail (same recipe as tests/golden/synth.py, CUDA generator)."""
    import math
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.empty(B, 1, CLIP_SAMPLES, device=dev)
    t = torch.arange(CLIP_SAMPLES, device=dev, dtype=torch.float32) / 22050.0
    chunk = 32
    for b0 in range(0, B, chunk):
        n = min(chunk, B - b0)
        s = 0.1 * torch.randn(n, CLIP_SAMPLES, device=dev, generator=g)
        f = torch.rand(n, 4, device=dev, generator=g) * 6900 + 100
        t0 = torch.rand(n, 4, device=dev, generator=g) * 48
        d = torch.rand(n, 4, device=dev, generator=g) * 10 + 1
        for j in range(4):
            m = (t[None] >= t0[:, j:j + 1]) & (t[None] < (t0[:, j:j + 1] + d[:, j:j + 1]))
            s += 0.5 * torch.sin(2 * math.pi * f[:, j:j + 1] * t[None]) * m
        x[b0:b0 + n, 0] = s
    for b in range(7, B, 8):
        cut = int((0.35 + 0.55 * ((b * 2654435761) % 1000) / 1000.0) * CLIP_SAMPLES)
        x[b, 0, cut:] = 0
    return x


def build_model(dev, dtype="bf16", deploy=True):
    import synth
    import yad_b200
    m = yad_b200.AudioDetectionNetwork(2, compute_dtype=dtype)
    skip = {"sm_anchors", "md_anchors", "lg_anchors", "taper_window"}
    layout = {k: tuple(v.shape) for k, v in m.state_dict().items()
              if k not in skip and "tfmr" not in k and "resampler" not in k}
    full = dict(m.state_dict())
    full.update(synth.synth_state_dict(layout, 42))
    m.load_state_dict(full)
    if deploy:
        m.inference()
    return m.eval().to(dev), full


def time_stages(model, x, reps=3):
    """Per-stage device time (ms) with CUDA events on the launching stream, outside the timed region."""
    eng = model._engine()
    B, _, L = x.shape
    model(x, combine_scales=True)          # make sure this shape's plan (workspaces) exists
    plan = eng._plan((B, L))
    out = {}

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(fn):
        best = []
        for _ in range(reps):
            a, b = ev(), ev()
            a.record(); r = fn(); b.record(); torch.cuda.synchronize()
            best.append(a.elapsed_time(b))
        return sum(best) / len(best), r

    import ctypes as C
    from yad_b200 import _lib
    T = eng.frames(L)
    mel = plan.get("mel")
    xc = x.contiguous()

    def fe_a():
        _lib.check(eng.lib.yad_frontend_mel_power(xc.data_ptr(), B, L, eng.rs_P, eng.rs_O, eng.rs_width, eng.rs_taps.data_ptr(),
                                                  eng.rs_base.data_ptr(), _lib.ptr(eng.rs_lane_map), eng.rs_window_len, eng.win.data_ptr(), eng.tw.data_ptr(),
                                                  eng.fb_val.data_ptr(), eng.fb_bin.data_ptr(), eng.fb_start.data_ptr(),
                                                  eng.fb_val.numel(), mel.data_ptr(), T, eng._stream()), "fe_a")
    out["frontend_mel_ms"], _ = timed(fe_a)
    out["frontend_ms"], xs = timed(lambda: eng.run_frontend(x, plan))

    def stem():
        if getattr(eng, "fused_stem", False):      # conv1 o conv2 as one convolution (+ border-column fix-up)
            cur = plan["c2"]
            fx = plan["fstem_cols"]
            xb = plan["xs_bf16"]
            _lib.check(eng.lib.yad_conv_stem_fused(xb.data_ptr(), xb.shape[2], B, 32, T, eng.fstem_w.data_ptr(), eng.fstem_bias.data_ptr(), cur.data_ptr(),
                                                   cur.shape[2], cur.shape[1], int(os.environ.get("YAD_STEM_NINT", "0")), eng._stream()), "stem_fused")
            _lib.check(eng.lib.yad_conv_stem_fused_fixup(xs.data_ptr(), B, 32, T, eng.fstem_wvar.data_ptr(), eng.fstem_bias.data_ptr(), fx[0],
                                                         fx[1], fx[2], cur.data_ptr(), cur.shape[2], cur.shape[1], eng._stream()), "stem_fixup")
            return
        c1 = plan["c1"]
        if eng.dtype == _lib.BF16:
            _lib.check(eng.lib.yad_conv_stem_tc(xs.data_ptr(), B, 32, T, eng.stem_w_tc.data_ptr(), c1.data_ptr(), c1.shape[2], c1.shape[1],
                                                eng._stream()), "stem_tc")
        else:
            _lib.check(eng.lib.yad_conv_stem(xs.data_ptr(), B, 32, T, eng.stem_w.data_ptr(), c1.data_ptr(), eng.dtype, eng._stream()), "stem")
    out["stem_ms"], _ = timed(stem)
    out["cnn_ms"], heads = timed(lambda: eng.run_cnn(xs, plan))
    L_res = -(-eng.rs_P * L // eng.rs_O)
    out["decode_ms"], preds = timed(lambda: eng.run_decode(heads, B, T, L_res))
    import yad_b200
    out["nms_ms"], _ = timed(lambda: yad_b200.nms_raw(preds, 0.1, 0.2, want_keep=False))
    return out


REF_CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def load_cpu_reference():
    """The CPU implementation of the path that the reference arm and `cpu_baseline` time: the LIVE reference (its own
    `modules.AudioDetectionNetwork` + `inference.process_model_outputs`, imported unmodified with a matplotlib stub) when a
    checkout is present (the dev container: /root/reference), else the oracle port (oracle/ref_port.py; the GPU box has no
    reference checkout - a Python reference does not travel).  Returns (kind, fn) with fn(x[B,1,L]) running forward +
    process_model_outputs on the synthetic weights of the GPU arm."""
    import synth
    _, full = build_model(torch.device("cpu"), deploy=False)
    sd = {k: v.cpu() for k, v in full.items()}
    ref = next((p for p in REF_CANDIDATES if os.path.exists(os.path.join(p, "modules", "_architecture.py"))), None)
    if ref is not None:
        try:
            import types
            mp = types.ModuleType("matplotlib"); mp.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules.setdefault("matplotlib", mp); sys.modules.setdefault("matplotlib.pyplot", mp.pyplot)
            sys.path.insert(0, ref)
            cwd = os.getcwd()
            os.chdir(ref)
            try:
                from modules import AudioDetectionNetwork as RefNet
                import inference as ref_inf
            finally:
                os.chdir(cwd)
            m = RefNet(2, config=os.path.join(ref, "config", "config.yaml"))
            live = dict(m.state_dict())
            live.update({k: v for k, v in sd.items() if k in live and live[k].shape == v.shape and k != "taper_window"})
            m.load_state_dict(live)
            m.eval()

            def run_live(x):
                out = m(x, combine_scales=True)
                try:
                    ref_inf.process_model_outputs(out, iou_threshold=0.1, conf_threshold=0.2)
                except (ValueError, RuntimeError):
                    pass
            return "reference", run_live
        except Exception as e:  # noqa: BLE001
            print(f"[bench] live reference at {ref} not importable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
    from oracle import ref_port as O

    def run_port(x):
        out = O.forward(x, sd, 2)
        try:
            O.process_model_outputs(out, 0.1, 0.2)
        except ValueError:
            pass
    return "port", run_port


def cpu_baseline(sample_clips=32, runs=9):
    """The reference path on the host cores (all threads) on a bounded sample of the GPU arm's workload, plus the B = 1 figure
    of BASELINE configs[0]."""
    import synth
    torch.set_num_threads(os.cpu_count() or 1)
    kind, fn = load_cpu_reference()
    x = synth.synth_clips(sample_clips, CLIP_SAMPLES, seed=1000, silence_tail_every=8)
    times, t1 = [], []
    with torch.no_grad():
        for i in range(runs + 1):
            t0 = time.perf_counter()
            fn(x)
            if i:
                times.append(time.perf_counter() - t0)
        for i in range(6):
            t0 = time.perf_counter()
            fn(x[:1])
            if i:
                t1.append(time.perf_counter() - t0)
    t = sorted(times)[len(times) // 2]
    tb1 = sorted(t1)[len(t1) // 2]
    return {"value": CLIP_SECONDS * sample_clips / t, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{sample_clips} synthetic 60 s clips, eval-mode train-form fp32 (the path inference.py runs), "
                      f"forward + process_model_outputs, median of {runs} runs, {t:.2f} s/run",
            "config0_b1": {"value": CLIP_SECONDS / tb1, "unit": "audio-s/s", "ms": 1e3 * tb1,
                           "what": "BASELINE configs[0]: one clip, batch = 1, forward + decode/NMS on the host cores, median of 5"}}, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    sample = 32
    import synth
    torch.set_num_threads(os.cpu_count() or 1)
    kind, fn = load_cpu_reference()
    x = synth.synth_clips(sample, CLIP_SAMPLES, seed=1000, silence_tail_every=8)
    budget = time.perf_counter() + 240
    with torch.no_grad():
        for _ in range(warm):                      # the requested warm-up (bounded by a quarter of the time budget)
            fn(x)
            if time.perf_counter() > budget - 180:
                break
        t0 = time.perf_counter(); done = 0
        for _ in range(steps):
            fn(x)
            done += 1
            if time.perf_counter() > budget:
                break
        dt = time.perf_counter() - t0
    v = CLIP_SECONDS * sample * done / dt
    what = ("the LIVE reference (modules.AudioDetectionNetwork + inference.process_model_outputs, unmodified)" if kind == "reference"
            else "oracle port (torch CPU fp32) of the reference path - no reference checkout on this machine")
    line = {"metric": METRIC, "value": v, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": done, "warmup": warm,
            "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "full pipeline: PCM -> log-mel/MFCC -> ResNet + RepBi-PAN -> decode -> NMS, "
                                   f"{sample} clips x 60 s per step (bounded sample of the GPU arm's 512-clip step; the metric is per clip)",
                       "clips_per_step": sample, "clip_seconds": CLIP_SECONDS, "num_classes": 2},
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": kind,
                             "sample": f"{sample} clips/step x {done} steps after {warm} warm-up steps, {what}"},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


TRAIN_METRIC = "audio-seconds/sec (train step: fwd + YOLO loss + bwd + Adam + EMA, DP gradient all-reduce)"
TRAIN_FLOP_PER_CLIP = 3 * 2 * 1514.2e6          # ~3x the train-form forward (1 514.2 nominal MMAC)


def cpu_baseline_train(sample_clips=4, runs=3):
    """The reference's train step (oracle port: torch CPU fp32 autograd, all host threads) on a bounded sample: train-mode
    forward (dropout 0: the oracle is deterministic) + loss + backward + Adam + EMA for `sample_clips` clips."""
    import synth
    from oracle import ref_port as O
    torch.set_num_threads(os.cpu_count() or 1)
    _, full = build_model(torch.device("cpu"), deploy=False)
    sd = {k: v.cpu().clone() for k, v in full.items()}
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and (k.endswith((".weight", ".bias")) or k.endswith("_anchors"))
             and "tfmr" not in k]
    x = synth.synth_clips(sample_clips, CLIP_SAMPLES, seed=2000, silence_tail_every=8)
    tg = synth.synth_targets(sample_clips, seed=11)
    cfg = dict(O.DEFAULT_CONFIG, dropout=0.0)
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    ema = {k: sd[k].clone() for k in names}
    times = []
    for it in range(runs + 1):
        t0 = time.perf_counter()
        with torch.enable_grad():
            for k in names:
                sd[k] = sd[k].detach().requires_grad_(True)
            loss, _ = O.detection_loss(O.forward_train(x, sd, 2, cfg), tg, O.DEFAULT_CONFIG["anchors"], 2)
            loss.backward()
        with torch.no_grad():
            ps = [sd[k] for k in names]
            O.adam_step(ps, [p.grad for p in ps], [m[k] for k in names], [v[k] for k in names], it + 1)
            O.ema_update([ema[k] for k in names], ps, it + 1)
        if it:
            times.append(time.perf_counter() - t0)
    t = sorted(times)[len(times) // 2]
    return {"value": CLIP_SECONDS * sample_clips / t, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_clips} synthetic 60 s clips per step, oracle port of the reference train step (train-mode forward, loss, "
                      f"autograd backward, Adam, EMA; torch CPU fp32), median of {runs} steps, {t:.2f} s/step"}


def measure_train(args, world, rank, local, dev, K, W, batch=None, e2e=True, lead_in=6):
    """BASELINE configs[4]: the data-parallel train step of pipeline/_trainer.py:94-108 timed on this rank set - forward, YOLO
    loss, backward, gradient all-reduce (bucketed behind the backward when N > 1), fused Adam + EMA.  Returns the numbers of the
    train line / the `train` block of the default line; assumes torch.distributed is initialised when world > 1."""
    import torch.distributed as dist
    import synth
    import yad_b200
    from yad_b200 import _lib, parallel
    cfg = yad_b200.default_config()
    tc = cfg["train_config"]
    B = int(batch or tc["batch_size"])                                      # 32 clips per GPU (config.yaml:59)
    torch.manual_seed(42)
    with torch.enable_grad():
        model, _ = build_model(dev, "bf16", deploy=False)
    model.train()
    model.train_graphs = not args.no_graphs       # frontend + forward and the backward replayed from CUDA graphs
    oc, ec = tc["optimizer_config"], tc["ema_config"]
    opt = yad_b200.FusedAdamEMA(model.parameters(), lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"], weight_decay=oc["weight_decay"],
                                ema_momentum=ec["momentum"], ema_N=ec["N"], use_ema=True)
    opt.overlap_allreduce(model)                  # gradient buckets are all-reduced while the rest of the backward runs
    eng = model._train_engine()
    hook = eng.on_bucket
    loss_fn = yad_b200.AudioDetectionLoss(cfg["anchors"], 2, sample_duration=cfg["sample_duration"], **tc["loss_config"])
    x = synth_clips_device(B, dev, seed=2000 + 7919 * rank)
    tg = synth.synth_targets(B, seed=11 + rank).to(dev)
    mode = {"ar": "overlap"}

    def step(xd, tgd):
        with torch.enable_grad():
            loss, met = loss_fn(model(xd), tgd)
            loss.backward()                       # bucket hooks fire inside (mode "overlap")
        if world > 1:
            if mode["ar"] == "overlap":
                opt.wait_allreduce()
            elif mode["ar"] == "serial":
                opt.allreduce_grads()
        opt.step()
        opt.zero_grad()
        return met

    def region(k, lead=lead_in):
        """k steps between CUDA events behind `lead` untimed lead-in steps; max over ranks (ms)."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect()
        gc.disable()
        for _ in range(lead):
            step(x, tg)
        e0.record()
        for _ in range(k):
            m_ = step(x, tg)
        e1.record()
        gc.enable()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return parallel.max_over_ranks(e0.elapsed_time(e1), device=dev), m_

    with torch.enable_grad():
        t_ramp = time.perf_counter()           # untimed clock ramp, see the inference workload
        ramp_s = float(os.environ.get("YAD_BENCH_RAMP_S", "1.5"))
        while True:
            # every step holds collectives (the gradient buckets): all ranks must run the SAME number of ramp steps, so rank 0's
            # clock decides and the decision is broadcast (a per-rank time test let one rank take one step more and deadlocked
            # the others in the next barrier - seen at N = 2)
            go = torch.tensor([1 if time.perf_counter() - t_ramp < ramp_s else 0], device=dev, dtype=torch.int32)
            if world > 1:
                dist.broadcast(go, 0)
            if int(go.item()) == 0:
                break
            step(x, tg)
            torch.cuda.synchronize()
        for _ in range(W):
            step(x, tg)
        torch.cuda.synchronize()
        n0 = _lib.launch_count
        met = step(x, tg)
        launches = _lib.launch_count - n0
        torch.cuda.nvtx.range_push("timed_train")
        regs = [region(K) for _ in range(3)]
        torch.cuda.nvtx.range_pop()
        ms = sorted(r[0] for r in regs)[1]
        met = regs[-1][1]
        out = {"clips_per_gpu": B, "ms_per_step": ms / K, "regions_ms_per_step": [r[0] / K for r in regs],
               "value": CLIP_SECONDS * B * world * K / (ms / 1e3), "unit": "audio-s/s", "dtype": "tf32", "launches_per_step": launches,
               "loss": met.get("aggregate_loss"), "allreduce": "none (1 GPU)"}
        if world > 1:
            # exposed all-reduce = step time with the overlapped all-reduce - step time with NO all-reduce (same kernels otherwise);
            # the serial form (backward, then one all-reduce of the whole arena: round 1's step) is timed for comparison
            mode["ar"] = "none"; eng.on_bucket = None
            ms_none = sorted(region(K)[0] for _ in range(3))[1]
            mode["ar"] = "serial"
            ms_serial = sorted(region(K)[0] for _ in range(3))[1]
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(); opt.allreduce_grads(); b_.record(); torch.cuda.synchronize()
            mode["ar"] = "overlap"; eng.on_bucket = hook
            out.update({"allreduce": "3 buckets behind the backward (36.6 / 8.4 / 3.5 MB fp32, NCCL AVG on its own stream)",
                        "ms_per_step_no_allreduce": ms_none / K, "ms_per_step_serial_allreduce": ms_serial / K,
                        "allreduce_ms": parallel.max_over_ranks(a.elapsed_time(b_), device=dev),
                        "exposed_allreduce_ms": (ms - ms_none) / K,
                        "exposed_allreduce_ms_serial": (ms_serial - ms_none) / K})
        if e2e:
            # end to end: pinned host PCM + targets -> device every step, loss metrics back (the loss's own 192-byte read)
            xh = torch.empty((B, 1, CLIP_SAMPLES), dtype=torch.float32, pin_memory=True); xh.copy_(x)
            th = tg.cpu().pin_memory()
            Ke = max(1, min(K, 5))
            step(xh.to(dev, non_blocking=True), th.to(dev, non_blocking=True))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(Ke):
                step(xh.to(dev, non_blocking=True), th.to(dev, non_blocking=True))
            e1.record()
            torch.cuda.synchronize()
            ms_e = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev)
            out["e2e"] = {"value": CLIP_SECONDS * B * world * Ke / (ms_e / 1e3), "unit": "audio-s/s",
                          "h2d_bytes_per_step": B * CLIP_SAMPLES * 4 + th.numel() * 4, "d2h_bytes_per_step": 3 * 8 * 8 + 3 * 4 * 4, "steps": Ke}
            del xh
        out["_step"] = lambda: step(x, tg)
    return out


def run_train(args):
    """`--workload train`: the train step as its own JSON line (the default line carries the same numbers in its `train` block)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(3, args.warmup), max(1, args.steps)
    t = measure_train(args, world, rank, local, dev, K, W, batch=args.batch if args.batch != 512 else None)
    sampler = ClockSampler(local)
    for _ in range(3):              # clocks under the same load, after every timed measurement (see the inference workload)
        for _ in range(min(K, 5)):
            t["_step"]()
        sampler.sample()
        torch.cuda.synchronize()
    if rank == 0:
        peaks = load_peaks()
        B = t["clips_per_gpu"]
        tfl = TRAIN_FLOP_PER_CLIP * B / (t["ms_per_step"] / 1e3) / 1e12
        line = {"metric": TRAIN_METRIC, "value": t["value"], "unit": "audio-s/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": t["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
                "data": "synthetic",
                "config": {"workload": f"train step (BASELINE configs[4]): {B} clips x 60 s per GPU, train-form net in train() mode "
                                       "(batch-stat BatchNorm, dropout 0.4), fp32 tensors with TF32 tcgen05 convolutions (cuDNN's default for fp32 training), "
                                       "YOLO loss, backward, fused Adam + EMA, forward / backward replayed from CUDA graphs; "
                                       "the 48.5 MB fp32 gradient arena is all-reduced in 3 buckets behind the backward when N > 1",
                           "clips_per_gpu": B, "clip_seconds": CLIP_SECONDS, "num_classes": 2,
                           "l2_policy": "per-step working set (169 MB PCM + 1.5 GB activations / gradients) is larger than the 126 MB L2",
                           "parallelism": f"data-parallel x{world}"},
                "e2e": t.get("e2e"), "gpu_launches": t["launches_per_step"] * K, "launches_per_step": t["launches_per_step"],
                "clocks": sampler.summary(),
                "roofline": {"kernel": "corr_tf32_kernel / wgrad_tf32_kernel (tcgen05 kind::tf32)", "bound": "tensor", "achieved": tfl,
                             "peak": peaks["bf16_tflops_sustained"] / 2, "unit": "TFLOP/s", "frac": tfl / (peaks["bf16_tflops_sustained"] / 2),
                             "traffic": None, "note": "whole-step useful conv FLOPs (3 x train-form forward) / step time; peak = half the "
                                                      "measured sustained bf16 rate (TF32 runs at half the bf16 MMA rate); at 32 clips per GPU "
                                                      "the step is bound by memory / launch latency of ~900 small kernels, not by the MMAs"},
                "train": {k: v for k, v in t.items() if not k.startswith("_") and k != "e2e"},
                "allreduce_ms": t.get("allreduce_ms", 0.0), "loss": t.get("loss"), "cpu_baseline": None}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_train()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="clips per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer: the headline metric (BASELINE configs[1]+[2]); train: the data-parallel train step (configs[4])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="train workload: launch kernel by kernel instead of replaying CUDA graphs")
    ap.add_argument("--e2e-chunk", type=int, default=32, help="clips per H2D chunk of the end-to-end pipeline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "train":
        return run_train(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    import yad_b200
    from yad_b200 import _lib, parallel, hostpipe
    from yad_b200.postprocess import segments_device

    # before any pinned allocation: this rank's CPUs = the NUMA node of its GPU (staging buffers first-touched node-locally)
    numa = hostpipe.bind_to_gpu_numa_node(local) if os.environ.get("YAD_BENCH_NUMA", "1") != "0" else {"bound": False, "why": "disabled"}

    W, K, B = max(3, args.warmup), max(1, args.steps), args.batch
    model, _ = build_model(dev, args.dtype, deploy=True)
    x = synth_clips_device(B, dev, seed=1000 + 7919 * rank)       # this rank's shard of the global batch

    def step(inp):
        preds = model(inp, combine_scales=True)
        # the device side of process_model_outputs: per-clip NMS + compaction into the reference's (segments, batch_idxs) layout
        return preds, segments_device(preds, 0.1, 0.2)

    sampler = ClockSampler(local)      # main-thread samples while enqueued work executes (warm-up, timed region, end-to-end loop)
    # Untimed ramp before the W warm-up steps: a GPU that has been idle (fresh box, first process) needs on the order of a
    # second of load to reach its full SM / memory clocks; with only W = 3 steps (12 ms) of warm-up the first bench on a fresh
    # box read 5.2 ms per step and every later one 3.9 ms.
    t_ramp = time.perf_counter()
    while time.perf_counter() - t_ramp < float(os.environ.get("YAD_BENCH_RAMP_S", "1.5")):
        step(x)
        torch.cuda.synchronize()
    for _ in range(W):
        step(x)
    torch.cuda.synchronize()
    n0 = _lib.launch_count
    step(x)
    launches_per_step = _lib.launch_count - n0
    torch.cuda.synchronize()

    # The K-step region is timed three times (each: barrier + synchronize, CUDA events around exactly K steps, synchronize +
    # barrier) and the MEDIAN region is reported; all three are in the JSON line ("regions_ms_per_step").  Reason: on the shared
    # boxes of this pool about one region in four reads 5-15 ms per step instead of 3.7 with unchanged per-stage times and clocks,
    # although the whole region is pre-queued on the device (below) - interference from outside the process (cf. what our own
    # nvidia-smi / NVML queries do to a running kernel stream, ClockSampler); one disturbed region must not decide the number.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    regions = []
    for _rep in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # the NVTX range (ncu --nvtx --nvtx-include "timed/" profiles exactly the timed steps) is opened BEFORE the start event:
        # its first call in a process loads the tools library - tens of milliseconds from a cold page cache
        torch.cuda.nvtx.range_push("timed")
        # no cyclic garbage collection inside the region (a generation-2 pass over this process's objects takes ~100 ms)
        gc.collect()
        gc.disable()
        # lead-in: 12 untimed steps are queued immediately before the start event.  The device has just been idle (synchronize,
        # a garbage collection of up to a few hundred ms) and drops its clocks within that time; the lead-in brings it back to
        # load clocks and gives the launching thread a ~45 ms head start, so the timed steps run back to back from a full
        # queue.  (Measured: without it the FIRST region of a process read 4.9-8.2 ms per step, the second and third 3.73; a
        # spin kernel as lead-in - queue full but device idle - did not cure it, full-load steps do.)
        for _ in range(12):
            step(x)
        e0.record()
        for _ in range(K):
            r = step(x)
        e1.record()
        gc.enable()
        torch.cuda.nvtx.range_pop()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        regions.append(parallel.max_over_ranks(e0.elapsed_time(e1), device=dev))
    ms = sorted(regions)[1]
    value = CLIP_SECONDS * B * world * K / (ms / 1e3)

    # ---- sustained: >= 2 s of back-to-back steps rotating over three DISTINCT 512-clip batches (so neither L2 nor anything else
    # sees a repeated input), one pair of events around the whole run.  The three burst regions above are 20 steps (~70 ms) each.
    # The same loop sized to BASELINE configs[3] (65 536 clips over all ranks, contiguous shards) gives the 64 k-clip line.
    xs_rot = [x] + [synth_clips_device(B, dev, seed=3000 + 7919 * rank + 101 * j) for j in (1, 2)]
    for xr in xs_rot:
        for _ in range(3):          # record -> replay -> graph capture for every buffer
            step(xr)
    torch.cuda.synchronize()

    def long_region(n_steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect(); gc.disable()
        for i in range(12):
            step(xs_rot[i % 3])
        a.record()
        for i in range(n_steps):
            step(xs_rot[i % 3])
        b.record()
        gc.enable()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return parallel.max_over_ranks(a.elapsed_time(b), device=dev)

    n_sus = max(K, int(2200.0 / (ms / K)) + 1)
    ms_sus = long_region(n_sus)
    sustained = {"steps": n_sus, "seconds": ms_sus / 1e3, "ms_per_step": ms_sus / n_sus,
                 "value": CLIP_SECONDS * B * world * n_sus / (ms_sus / 1e3), "unit": "audio-s/s",
                 "inputs": "three distinct 512-clip batches per GPU in rotation (8.1 GB of PCM), one CUDA-event pair around the run"}
    n64 = -(-65536 // (B * world))
    ms64 = long_region(n64)
    config3 = {"clips": n64 * B * world, "steps_per_gpu": n64, "seconds": ms64 / 1e3,
               "value": CLIP_SECONDS * n64 * B * world / (ms64 / 1e3), "unit": "audio-s/s",
               "what": "BASELINE configs[3]: 65 536 clips, contiguous shards over the ranks, processed in 512-clip steps per GPU "
                       "(rotating over three resident synthetic batches: 347 GB of distinct PCM does not fit), device-timed, max over ranks"}
    del xs_rot[1:]

    # ---- parity spot check of the benchmarked path (this process, these kernels, B = 512): 4 clips of the big batch's output vs
    # the same clips run as a batch of 4 - bitwise (batch position / CTA assignment / graph replay must not change a result),
    # and their segments
    preds_big, (seg_big, bidx_big, tot_big) = step(x)
    pick = [0, B // 3, B - 149 if B > 149 else B // 2, B - 1]
    small = model(x[pick].contiguous(), combine_scales=True)
    spot_ok = bool(torch.equal(small, preds_big[pick]))
    finite = bool(torch.isfinite(preds_big).all().item())
    nseg = int(tot_big.item())
    try:
        seg_s, bidx_s = yad_b200.process_model_outputs(small, 0.1, 0.2)
        bb = bidx_big[:nseg]
        ref_rows = torch.cat([seg_big[:nseg][bb == p_] for p_ in pick])
        spot_ok = spot_ok and bool(torch.equal(ref_rows, seg_s))
    except ValueError:
        pass
    parity_spot = {"status": "ok" if (spot_ok and finite) else "MISMATCH", "clips": pick, "segments_total": nseg,
                   "what": "preds of 4 clips inside the 512-clip batch == the same clips as a batch of 4, bitwise; segments equal; all finite"}
    del preds_big, small

    # ---- BASELINE configs[1]: the frontend alone at B = 256 (PCM f32 -> x_spectral), device-timed
    fe256 = None
    if B >= 256:
        eng = model._engine()
        x256 = x[:256]
        plan256 = eng._plan((256, CLIP_SAMPLES))
        for _ in range(3):
            eng.run_frontend(x256, plan256)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(6):
            eng.run_frontend(x256, plan256)
        a.record()
        for _ in range(K):
            eng.run_frontend(x256, plan256)
        b.record()
        torch.cuda.synchronize()
        t256 = a.elapsed_time(b) / K
        fe256 = {"clips": 256, "ms": t256, "value": CLIP_SECONDS * 256 / (t256 / 1e3), "unit": "audio-s/s",
                 "achieved_gbs": FRONTEND_BYTES_PER_CLIP * 256 / (t256 / 1e3) / 1e9,
                 "what": "BASELINE configs[1]: resample + log-mel + MFCC + dB + standardise, 256 clips, inputs resident (1.35 GB > L2)"}

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    xh = torch.empty((B, 1, CLIP_SAMPLES), dtype=torch.float32, pin_memory=True)
    xh.copy_(x)
    def e2e_step():
        # public host-buffer API: chunked, double-buffered H2D copy overlapped with the forward, one D2H of the segments
        return yad_b200.run_host_batch(model, xh, 0.1, 0.2, chunk=args.e2e_chunk)
    e2e_step()
    torch.cuda.synchronize()
    Ke = max(1, min(K, 5))
    if world > 1:
        dist.barrier()
    e0.record()
    d2h = 0
    for _ in range(Ke):
        seg, bidx = e2e_step()
        if seg is not None:
            d2h = seg.numel() * 4 + bidx.numel() * 8 + 8
    e1.record()
    torch.cuda.synchronize()
    ms_e = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev)
    e2e = {"value": CLIP_SECONDS * B * world * Ke / (ms_e / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": B * CLIP_SAMPLES * 4,
           "d2h_bytes_per_step": d2h, "steps": Ke}
    # the ceiling of that number: the raw pinned cudaMemcpyAsync rate of the same bytes in the same chunks, all ranks concurrently
    if world > 1:
        dist.barrier()
    nchunks = -(-B // args.e2e_chunk)
    ceil_gbs = hostpipe.h2d_ceiling_gbs(args.e2e_chunk * CLIP_SAMPLES * 4, nchunks, dev)
    ceil_min = -parallel.max_over_ranks(-ceil_gbs, device=dev)          # the slowest rank's link sets the max-over-ranks time
    e2e_gbs = B * CLIP_SAMPLES * 4 * Ke / (ms_e / 1e3) / 1e9
    # ... and of the very same host tensor (2.7 GB of distinct host memory per rank instead of one 170 MB buffer copied 16 times: with
    # several GPUs pulling at once the small buffer can be served from the host's last-level cache, the batch cannot)
    if world > 1:
        dist.barrier()
    ceil_d = hostpipe.h2d_ceiling_tensor_gbs(xh, args.e2e_chunk, dev)
    ceil_d_min = -parallel.max_over_ranks(-ceil_d, device=dev)
    e2e.update({"h2d_gbs_per_gpu": e2e_gbs, "h2d_ceiling_gbs": ceil_min, "frac_of_ceiling": e2e_gbs / ceil_min if ceil_min else None,
                "h2d_ceiling_same_tensor_gbs": ceil_d_min, "frac_of_ceiling_same_tensor": e2e_gbs / ceil_d_min if ceil_d_min else None,
                "numa": numa, "ceiling": f"{nchunks} back-to-back pinned cudaMemcpyAsync of {args.e2e_chunk} clips each on one stream, CUDA events, "
                                         "best of 3, all ranks concurrently, min over ranks"})
    # the same call with 16-bit PCM host buffers (the sample format of audio files; x / 32768 inside the frontend kernel): half
    # the PCIe bytes.  Reported next to `e2e`, not instead of it - the reference-facing API takes fp32.
    xi = torch.empty((B, 1, CLIP_SAMPLES), dtype=torch.int16, pin_memory=True)
    xi.copy_((x.clamp(-1, 1) * 32767).round().to(torch.int16))
    yad_b200.run_host_batch(model, xi, 0.1, 0.2, chunk=args.e2e_chunk)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(Ke):
        yad_b200.run_host_batch(model, xi, 0.1, 0.2, chunk=args.e2e_chunk)
    e1.record()
    torch.cuda.synchronize()
    ms_i = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev)
    e2e_i16 = {"value": CLIP_SECONDS * B * world * Ke / (ms_i / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": B * CLIP_SAMPLES * 2,
               "steps": Ke, "input": "int16 PCM (same clips quantised to 16 bit)"}
    del xi
    # Clocks / throttle reasons under exactly the benchmark's load: every timed measurement above is finished; the K steps and one
    # end-to-end step are run once more, untimed, and sampled while they execute.  No monitoring query (not even nvmlInit) runs
    # before this point: queries before or inside a timed region were measured to corrupt it (see ClockSampler).
    for _ in range(3):
        for _ in range(K):
            step(x)
        sampler.sample()
        torch.cuda.synchronize()
    e2e_step()
    sampler.sample()
    torch.cuda.synchronize()
    del xh

    # ---- the train step (BASELINE configs[4]) in the same line, so that the driver's 1 -> N scaling run carries the train curve:
    # 32 clips per GPU, gradient buckets all-reduced behind the backward (every rank takes part: NCCL)
    train_block = None
    if os.environ.get("YAD_BENCH_TRAIN", "1") != "0":
        t = measure_train(args, world, rank, local, dev, K=max(5, min(K, 10)), W=3, e2e=False)
        train_block = {k: v for k, v in t.items() if not k.startswith("_")}
        train_block["metric"] = TRAIN_METRIC
        del t
        torch.set_grad_enabled(False)

    if rank == 0:
        peaks = load_peaks()
        st = time_stages(model, x)
        conv_ms = st["cnn_ms"]
        fe_ms = st["frontend_mel_ms"]
        # the dominant kernel of the step is frontend_mel_kernel (one launch per step; see profiles/ launch list)
        ach = MEL_KERNEL_BYTES_PER_CLIP * B / (fe_ms / 1e3) / 1e9
        traffic_clip, traffic_src = load_mel_kernel_traffic()
        roof = {"kernel": "frontend_mel_kernel", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic_clip * B if traffic_clip else None, "traffic_source": traffic_src,
                "peak_source": peaks["source"],
                "algorithmic_bytes_per_launch": MEL_KERNEL_BYTES_PER_CLIP * B, "launch_ms": fe_ms,
                "note": "achieved = (PCM read + mel write) / CUDA-event time of the launch; DRAM traffic = algorithmic bytes; the kernel "
                        "is bound by the shared-memory data pipe, not by HBM (ncu r02: l1tex__data_pipe_lsu_wavefronts 78 % of peak, 4.7 k "
                        "wavefronts and 12.9 k warp instructions per 8-frame group, FMA pipe 52 %, barrier stalls 25 % of the warp "
                        "samples) - see DESIGN.md and profiles/r02_ncu_frontend_mel_b512_v3_by_line.txt"}
        fe_all = st["frontend_ms"]             # stage A + stage B: the whole frontend's bytes over the whole frontend's time
        roof["frontend_hbm"] = {"achieved_gbs": FRONTEND_BYTES_PER_CLIP * B / (fe_all / 1e3) / 1e9,
                                "frac": FRONTEND_BYTES_PER_CLIP * B / (fe_all / 1e3) / 1e9 / peaks["hbm_gbs"], "ms": fe_all}
        fused = getattr(model._engine(), "fused_stem", False)
        flop = CNN_FLOP_EXECUTED_PER_CLIP if fused else CNN_FLOP_PER_CLIP
        roof["cnn_tensor"] = {"achieved_tflops": flop * B / (conv_ms / 1e3) / 1e12,
                              "frac": flop * B / (conv_ms / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
                              "reference_graph_tflops": CNN_FLOP_PER_CLIP * B / (conv_ms / 1e3) / 1e12,
                              "note": "achieved = FLOPs this implementation executes (stem conv1 o conv2 composed: 1.70 instead of "
                                      "2.30 GFLOP per clip) / CNN stage time; reference_graph_tflops credits the reference's FLOPs"}
        line = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "regions_ms_per_step": [r_ / K for r_ in regions], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype,
                "data": "synthetic",
                "config": {"workload": f"full pipeline, deploy-form (reparameterised) net, {B} clips x 60 s per GPU per step: PCM f32 -> "
                                       "fused resample/log-mel/MFCC frontend -> fused stem (conv1 o conv2) + ResNet-18 + RepBi-PAN (tcgen05 implicit GEMM) -> "
                                       "anchor decode -> per-clip segment NMS (BASELINE configs[1]+[2] chained)",
                           "clips_per_gpu": B, "clip_seconds": CLIP_SECONDS, "num_classes": 2,
                           "l2_policy": "inputs (2.7 GB PCM + 0.8 GB activations per step) are larger than the 126 MB L2",
                           "parallelism": f"clip-sharded x{world}, no data-path collective",
                           "timing": "median of three identical regions (all in regions_ms_per_step), each: CUDA events around exactly K steps, "
                                     "barrier + synchronize on both sides; 12 untimed lead-in steps are queued right before the "
                                     "start event (device back at load clocks after the idle gap, timed steps run from a full "
                                     "queue); gc disabled inside; clocks sampled after all timed measurements during untimed repeats"},
                "e2e": e2e, "e2e_int16": e2e_i16, "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
                "clocks": sampler.summary(), "roofline": roof, "stages_ms": st, "sustained": sustained,
                "config3_64k_clips": config3, "config1_frontend_b256": fe256, "parity_spot": parity_spot, "train": train_block}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"], _ = cpu_baseline()
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
