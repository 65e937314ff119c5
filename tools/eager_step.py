"""Times the 512-clip inference step launched kernel by kernel (no CUDA graph) - the path a caller with ever-changing input
tensors takes.  Used to measure programmatic dependent launch (YAD_PDL=0/1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
os.environ.setdefault("YAD_INFER_GRAPHS", "0")
import torch
import bench
import yad_b200
from yad_b200.postprocess import segments_device
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
model, _ = bench.build_model(dev, "bf16", deploy=True)
x = bench.synth_clips_device(B, dev, 1000)
def step():
    p = model(x, combine_scales=True)
    return segments_device(p, 0.1, 0.2)
for _ in range(30):
    step()
torch.cuda.synchronize()
res = []
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10):
        step()
    a.record()
    for _ in range(20):
        step()
    b.record()
    torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / 20)
print("YAD_PDL=%s graphs=%s B=%d ms/step %s" % (os.environ.get("YAD_PDL", "1"), os.environ["YAD_INFER_GRAPHS"], B, ["%.3f" % r for r in res]))
