"""Raw pinned host -> device copy rate of a 2.7 GB batch in chunks, over 1 / 2 / 3 copy streams and several chunk sizes
(is there anything to gain from more than one copy in flight?).  Usage: python tools/h2d_streams.py"""
import torch
dev = torch.device("cuda", 0)
CL = 1323000
N = 512
x = torch.empty((N, 1, CL), dtype=torch.float32, pin_memory=True)
x.zero_()
for chunk in (16, 32, 64, 128):
    for ns in (1, 2, 3):
        streams = [torch.cuda.Stream(dev) for _ in range(ns)]
        dst = [torch.empty((chunk, 1, CL), device=dev) for _ in range(2 * ns)]
        best = 0.0
        for rep in range(3):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for st in streams:
                st.wait_stream(torch.cuda.current_stream(dev))
            for i, s0 in enumerate(range(0, N, chunk)):
                st = streams[i % ns]
                with torch.cuda.stream(st):
                    dst[i % (2 * ns)].copy_(x[s0:s0 + chunk], non_blocking=True)
            for st in streams:
                torch.cuda.current_stream(dev).wait_stream(st)
            b.record()
            torch.cuda.synchronize()
            best = max(best, x.numel() * 4 / (a.elapsed_time(b) / 1e3) / 1e9)
        print(f"chunk {chunk:4d} clips, {ns} stream(s): {best:6.2f} GB/s")
