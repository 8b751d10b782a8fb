"""One eager step of the bench workload (64 pairs of C2) between cudaProfilerStart/Stop, after a warm-up step: run under
`ncu --profile-from-start off --set full ...` to capture exactly the kernels of one step.  Usage: prof_step.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lk_optical_flow_method_b200.pipeline import PairPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
frames, pts = bench.make_inputs(B, 0, dev)
pipe = PairPipeline(B, bench.H, bench.W, bench.NPTS, bench.LK, bench.GFTT, bench.REFRESH)
pipe.overlap = False                      # serial launches: one kernel at a time under the profiler
pipe.prime(frames[0])
for i in range(2):
    pipe.set_points(pts)
    if i == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    pipe.step(frames[1:])
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
