"""Minimal driver for ncu: builds the bench model, runs `--warm` untimed steps and `--steps` steps of the hot path.
Used for the launch list / --set full captures under profiles/ (numbers printed under ncu are never bench values)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    import yad_b200
    from yad_b200 import _lib
    torch.set_grad_enabled(False)
    dev = torch.device("cuda", 0)
    model, _ = bench.build_model(dev, a.dtype, deploy=True)
    x = bench.synth_clips_device(a.batch, dev, 1000)
    for i in range(a.warm + a.steps):
        n0 = _lib.launch_count
        preds = model(x, combine_scales=True)
        yad_b200.nms_raw(preds, 0.1, 0.2, want_keep=False)
        torch.cuda.synchronize()
        if i == 0:
            print("launches per step:", _lib.launch_count - n0)
    print("done")


if __name__ == "__main__":
    main()
