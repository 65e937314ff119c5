"""Fused neck kernel vs the layer-by-layer neck on the same backbone maps: heads and (debug build of the program) the
intermediate planes p4 / bic3 / p3 / bic2 dumped from shared memory."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import bench
import yad_b200
from yad_b200.neck_fused import FusedNeck
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1323000
model, _ = bench.build_model(dev, "bf16", deploy=True)
x = bench.synth_clips_device(B, dev, 1000)[:, :, :L].contiguous()
eng = model._engine()
plan = eng._plan((B, L))
xs = eng.run_frontend(x, plan)
eng.fused_neck = False
heads_ref = [h.clone() for h in eng.run_cnn(xs, plan)]
torch.cuda.synchronize()
# the flat backbone maps of the plan
fm = [plan[f"s{li}.1.y"] for li in range(4)]
geoms = [(f.shape[2] - 1 if f.shape[2] > 1 else 1, f.shape[1] - 1) for f in fm]
print("geoms", geoms, [tuple(f.shape) for f in fm])
fn = FusedNeck(eng, [g[0] for g in geoms], [g[1] for g in geoms], [f.shape[3] for f in fm], debug=True, G=1)   # one clip per pass: the dumps are per clip (two clips per pass are pinned bitwise to this by tests/test_gpu_parity.py)
print("program: ops", len(fn.ops), "kbs", len(fn.kbs), "pool", fn.pool_bytes, "slots", fn.n_slots)
heads = [torch.full_like(h, float("nan")) for h in heads_ref]
dbg = torch.zeros(B * fn.dump_elems, dtype=torch.bfloat16, device=dev)
fn.run(eng.lib, fm, heads, eng._stream(), dbg=dbg)
torch.cuda.synchronize()
dbg = dbg.view(B, -1).float()
def plane(name, W):
    off, rows, lvl = fn.dumps[name]
    return dbg[:, off:off + rows * 64].view(B, rows, 64)[:, 1:1 + W, :]
refs = {"p4": plan["cat_n4"][:, 0, :, :128].float(), "b3": plan["b3"][:, 0].float(), "p3": plan["cat_n3"][:, 0, :, :128].float(),
        "b2": plan["b2"][:, 0].float()}
def rep(name, got, ref):
    d = (got - ref).abs()
    print(f"{name:8s} max|ref| {ref.abs().max():.3f}  max diff {d.max():.4f}  mean diff {d.mean():.5f}  nan {int(torch.isnan(got).sum())}")
W4, W3, W2 = geoms[3][1], geoms[2][1], geoms[1][1]
rep("a1", plane("a1", W4), plan["sp.a1"][:, 0].float())
rep("y", plane("y", W4), plan["sp.cat7"][:, 0, :, 64:128].float())
rep("a2", plane("a2", W4), plan["sp.a2"][:, 0].float())
rep("a3", plane("a3", W4), plan["sp.cat5"][:, 0, :, 0:64].float())
rep("m1", plane("m1", W4), plan["sp.cat5"][:, 0, :, 64:128].float())
rep("m3", plane("m3", W4), plan["sp.cat5"][:, 0, :, 192:256].float())
rep("b1", plane("b1", W4), plan["sp.c5"][:, 0].float())
g_, r_ = plane("a1", W4), plan["sp.a1"][:, 0].float()
print("a1 got[0,:3,:6]", g_[0, :3, :6].tolist()); print("a1 ref[0,:3,:6]", r_[0, :3, :6].tolist())
rep("p4", torch.cat([plane("p4_0", W4), plane("p4_1", W4)], -1), refs["p4"])
rep("b3[:64]", plane("b3_0", W3), refs["b3"][..., :64])
rep("p3", torch.cat([plane("p3_0", W3), plane("p3_1", W3)], -1), refs["p3"])
rep("b2", torch.cat([plane("b2_0", W2), plane("b2_1", W2)], -1), refs["b2"])
for i, (h, r) in enumerate(zip(heads, heads_ref)):
    rep(f"head{i}", h[..., :15], r[..., :15])
# guard / halo rows of a dumped plane must be zero
off, rows, _ = fn.dumps["p3_0"]
pl = dbg[:, off:off + rows * 64].view(B, rows, 64)
print("p3_0 guard row max", pl[:, 0].abs().max().item(), "halo row max", pl[:, 1 + W3].abs().max().item())
