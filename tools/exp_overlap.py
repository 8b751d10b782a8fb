"""Experiment: does the stencil pass (process_img + pyramid of the NEXT batch, HBM-bound) hide behind the LK launch of the
current batch (issue-bound) when it runs on a high-priority stream?  Prints serial and concurrent times of the two."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lk_optical_flow_method_b200.pipeline import PairPipeline

B = 64
dev = torch.device("cuda", 0)
frames, pts = bench.make_inputs(B, 0, dev)
pipe = PairPipeline(B, bench.H, bench.W, bench.NPTS, bench.LK, bench.GFTT, 0)
pipe.prime(frames[0]); pipe.set_points(pts)
pipe.stage_process_pyramid(frames[1:])
other = PairPipeline(B, bench.H, bench.W, bench.NPTS, bench.LK, bench.GFTT, 0)      # receives the "next batch" stencils
other.prime(frames[0])
hi = torch.cuda.Stream(priority=-1)
lo = torch.cuda.Stream(priority=0)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def serial():
    other.stage_process_pyramid(frames[1:])
    pipe.stage_track()


def concurrent(stencil_stream):
    main = torch.cuda.current_stream()
    f, j = torch.cuda.Event(), torch.cuda.Event()
    f.record(main)
    with torch.cuda.stream(stencil_stream):
        stencil_stream.wait_event(f)
        other.stage_process_pyramid(frames[1:])
        j.record(stencil_stream)
    pipe.stage_track()
    main.wait_event(j)


print("stencils alone  %.1f us" % timed(lambda: other.stage_process_pyramid(frames[1:])))
print("LK alone        %.1f us" % timed(pipe.stage_track))
print("serial          %.1f us" % timed(serial))
print("concurrent (stencils on the high-priority stream) %.1f us" % timed(lambda: concurrent(hi)))
print("concurrent (stencils on a normal stream)          %.1f us" % timed(lambda: concurrent(lo)))
