"""Times the train step (forward + loss + backward + fused Adam/EMA) at B clips of 60 s on one GPU; prints a stage split."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import yad_b200, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
cfg = yad_b200.default_config()
torch.manual_seed(42)
m = yad_b200.AudioDetectionNetwork(2, config=cfg, train_dtype=(sys.argv[3] if len(sys.argv) > 3 else "tf32")).to(dev).train()
m.train_graphs = len(sys.argv) > 4 and sys.argv[4] == "graph"
opt = yad_b200.FusedAdamEMA(m.parameters(), lr=1e-3, weight_decay=0.002, ema_momentum=0.002, use_ema=True)
loss_fn = yad_b200.AudioDetectionLoss(cfg["anchors"], 2, sample_duration=60, **cfg["train_config"]["loss_config"])
x = (torch.randn(B, 1, 1323000, device=dev) * 0.1)
tg = synth.synth_targets(B, seed=11).to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(steps + 2):
    e = [ev() for _ in range(5)]
    import time as _t
    torch.cuda.synchronize()
    c0 = _t.perf_counter()
    e[0].record()
    with torch.enable_grad():
        preds = m(x)
        c1 = _t.perf_counter()
        e[1].record()
        loss, met = loss_fn(preds, tg)
        c2 = _t.perf_counter()
        e[2].record()
        loss.backward()
        c3 = _t.perf_counter()
        e[3].record()
    print(f"   cpu: fwd {1e3*(c1-c0):.2f} loss {1e3*(c2-c1):.2f} bwd {1e3*(c3-c2):.2f} ms")
    opt.step(); opt.zero_grad()
    e[4].record()
    torch.cuda.synchronize()
    print(f"it {it}: loss {float(loss):.4f}  fwd {e[0].elapsed_time(e[1]):.2f}  loss {e[1].elapsed_time(e[2]):.2f}  "
          f"bwd {e[2].elapsed_time(e[3]):.2f}  opt {e[3].elapsed_time(e[4]):.2f}  total {e[0].elapsed_time(e[4]):.2f} ms", flush=True)
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)
