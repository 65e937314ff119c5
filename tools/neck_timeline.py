"""Per-op timeline of the fused neck kernel (CTA 0, its first clip): clock64 stamps of the MMA warp (op start, last MMA issued) and of
the epilogue (accumulator complete, op done).  Usage: python tools/neck_timeline.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import bench
import yad_b200
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ITER = int(sys.argv[2]) if len(sys.argv) > 2 else 0       # which clip of CTA 0: 0 = first (cold), 1 = second (steady state)
os.environ["YAD_INFER_GRAPHS"] = "0"
model, _ = bench.build_model(dev, "bf16", deploy=True)
x = bench.synth_clips_device(B, dev, 1000)
eng = model._engine()
for _ in range(2):
    model(x, combine_scales=True)
fn = [v for v in eng._fused_necks.values() if v is not None][0]
n = len(fn.ops)
buf = torch.zeros(8 * n, dtype=torch.int64, device=dev)
eng.lib.yad_neck_fused_set_timeline_iter(ITER)
eng.lib.yad_neck_fused_set_timeline(buf.data_ptr())
model(x, combine_scales=True)
torch.cuda.synchronize()
eng.lib.yad_neck_fused_set_timeline(0)
eng.lib.yad_neck_fused_set_timeline_iter(0)
t = buf.view(n, 8).cpu().numpy()
t0 = min(v for v in t[:, :6].reshape(-1) if v > 0)
names = {0: "CONV", 1: "POOLS", 2: "PAIRAVG", 3: "UP2", 4: "DEINT", 5: "DUMP"}
print(f"{'op':>3} {'type':7} {'N':>4} {'nkb':>4} {'n_mt':>4} {'src':>4} | {'start':>7} {'issued':>7} {'acc':>7} {'done':>7} {'arr_w2':>7} {'arr_max':>7} | {'gap':>6} {'wfull':>6} {'issue':>6} {'drain':>6} {'epi':>6} {'total':>6}  (cycles; gap = previous op's last arrival -> MMA warp start)")
prev_done = 0
prev_arr = -1
for i, op in enumerate(fn.ops):
    a, b, c, d, e, f = [int(v - t0) if v > 0 else -1 for v in t[i][:6]]
    if op[0] == 0:
        print(f"{i:3d} {names[op[0]]:7} {op[2]:4d} {op[4]:4d} {op[1]:4d} {op[14]:4d} | {a:7d} {b:7d} {c:7d} {d:7d} {e:7d} {f:7d} | {a - prev_arr if prev_arr >= 0 else -1:6d} {int(t[i][6]):6d} {b - a:6d} {c - b:6d} {d - c:6d} {d - prev_done:6d}")
    else:
        print(f"{i:3d} {names[op[0]]:7} {'':4} {'':4} {'':4} {'':4} | {'':7} {'':7} {'':7} {d:7d} {e:7d} {f:7d} | {'':6} {'':6} {'':6} {'':6} {'':6} {d - prev_done:6d}")
    prev_done = d
    prev_arr = f
print("clip total cycles:", prev_done)

