"""Timings of the BASELINE.json configs that are parity-test cases rather than bench lines (C3, C4, C5), with the
reference's CPU implementation (cv2) timed beside them on the same inputs.  Device-resident inputs, CUDA events,
back-to-back repetitions.  Usage: python tools/bench_configs.py [c3] [c4] [c5]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lk_optical_flow_method_b200 import ops, synth, vp as V


def gpu_time(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def cpu_time(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(t))


def frames(W, H, n, seed=0, t0=3):
    tex = torch.from_numpy(synth.make_texture(W, H, seed)).cuda()
    return torch.stack([synth.zoom_frame(tex, W, H, t0 + i) for i in range(n)])


def c3():
    import cv2
    W, H, N, B = 3840, 2160, 10000, 4
    fr = frames(W, H, B + 1)
    gray = ops.process_img(fr)
    lk = dict(winSize=(31, 31), maxLevel=3, criteria=(3, 10, 0.03))
    nl = ops.effective_levels(H, W, lk["winSize"], lk["maxLevel"])
    cor, cnt = ops.gftt_detect(gray[:B], None, N, 0.001, 3, 3, N)
    pyr = ops.build_pyramid(gray, nl)
    prev, nxt = [p[:B] for p in pyr], [p[1:] for p in pyr]
    params = ops.lk_params_struct(**lk)
    out = {"config": "C3 3840x2160, 10000 features, 31x31, 4 levels", "pairs_per_launch": B, "corners_found": cnt.cpu().tolist()}
    out["gpu_ms_per_pair"] = {
        "process_img": gpu_time(lambda: ops.process_img(fr[1:], out=gray[1:])) / B,
        "pyramid": gpu_time(lambda: ops.build_pyramid(gray[1:], nl)) / B,
        "gftt": gpu_time(lambda: ops.gftt_detect(gray[:B], None, N, 0.001, 3, 3, N)) / B,
        "pyrlk": gpu_time(lambda: ops.pyrlk(prev, nxt, cor, params, counts=cnt)) / B}
    g0, g1 = gray[0].cpu().numpy(), gray[1].cpu().numpy()
    f1 = fr[1].cpu().numpy()
    p0 = cor[0, :int(cnt[0])].cpu().numpy().reshape(-1, 1, 2)
    cv2.setNumThreads(os.cpu_count())
    out["cv2_ms_per_pair"] = {
        "process_img": cpu_time(lambda: cv2.GaussianBlur(cv2.cvtColor(f1, cv2.COLOR_BGR2GRAY), (3, 3), 0)),
        "gftt": cpu_time(lambda: cv2.goodFeaturesToTrack(g0, mask=None, maxCorners=N, qualityLevel=0.001, minDistance=3, blockSize=3)),
        "pyrlk (incl. pyramids)": cpu_time(lambda: cv2.calcOpticalFlowPyrLK(g0, g1, p0, None, **lk))}
    out["cv2_threads"] = cv2.getNumThreads()
    return out


def c4():
    import cv2
    W, H = 1920, 1080
    fr = frames(W, H, 2)
    gray = ops.process_img(fr)
    lk = dict(winSize=(5, 5), maxLevel=2, criteria=(3, 10, 0.03))
    nl = ops.effective_levels(H, W, lk["winSize"], lk["maxLevel"])
    pyr = ops.build_pyramid(gray, nl)
    prev, nxt = [p[:1] for p in pyr], [p[1:] for p in pyr]
    ys, xs = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
    pts = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1).float()[None].contiguous()
    params = ops.lk_params_struct(**lk)
    res = ops.pyrlk(prev, nxt, pts, params, want_iters=True)
    it = float(res[3].float().mean().item())
    ms = gpu_time(lambda: ops.pyrlk(prev, nxt, pts, params), reps=5)          # grid detected on the device -> k_lk_dense
    buf = ops.pyrlk_dense(prev, nxt, params)
    ms_dense = gpu_time(lambda: ops.pyrlk_dense(prev, nxt, params, out=buf), reps=5)
    rev = pts.flip(1).contiguous()                                             # not the raster grid -> k_lk_track_small
    ms_small = gpu_time(lambda: ops.pyrlk(prev, nxt, rev, params), reps=3)
    one = ops.lk_params_struct(winSize=(5, 5), maxLevel=2, criteria=(1, 1, 0.0))
    ms_one = gpu_time(lambda: ops.pyrlk_dense(prev, nxt, one, out=buf), reps=5)
    out = {"config": "C4 dense 1920x1080, 5x5, 3 levels, criteria (EPS|COUNT,10,0.03)", "features": W * H,
           "gpu_ms": ms, "gpu_ms_dense_entry": ms_dense, "gpu_ms_thread_per_feature_kernel": ms_small,
           "gpu_ms_one_iteration_variant": ms_one, "gpu_ns_per_feature": 1e6 * ms / (W * H), "iterations_per_feature": it,
           "status_ok_share": float(res[1].float().mean().item())}
    # op model of SURVEY 8d: w h (27 L + 12 it + 9)
    out["gop"] = W * H * 25 * (27 * nl + 12 * it + 9) / 1e9
    out["achieved_tops"] = out["gop"] / ms
    g0, g1 = gray[0].cpu().numpy(), gray[1].cpu().numpy()
    p0 = pts[0].cpu().numpy().reshape(-1, 1, 2)
    cv2.setNumThreads(os.cpu_count())
    out["cv2_ms"] = cpu_time(lambda: cv2.calcOpticalFlowPyrLK(g0, g1, p0, None, **lk), reps=2)
    out["cv2_threads"] = cv2.getNumThreads()
    return out


def c5():
    W, H, n_clips, n_frames = 1920, 1080, 64, 12
    texs = [torch.from_numpy(synth.make_texture(W, H, c % 4)).cuda() for c in range(4)]       # 4 textures, 64 jittered clips
    def frame_batch(t):
        out = []
        for c in range(n_clips):
            rng = np.random.default_rng(c)
            foe = (0.48 * (1 + rng.uniform(-0.05, 0.05)), 0.60 * (1 + rng.uniform(-0.05, 0.05)))
            out.append(synth.zoom_frame(texs[c % 4], W, H, t, rate=float(rng.uniform(1.010, 1.020)), foe=foe))
        return torch.stack(out)
    batches = [frame_batch(t) for t in range(n_frames)]
    r = V.ClipBatchRunner(n_clips, (H, W), variant="VP", expected_frames=4 * n_frames, TP_NUM=60)
    r.start(batches[0])
    buf = torch.empty_like(batches[0])
    for fr in batches[1:3]:
        buf.copy_(fr); r.step_graph(buf)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for fr in batches[3:]:
        buf.copy_(fr)
        a.record(); r.step_graph(buf); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    steps = len(batches) - 3
    rows = r.rows_host()
    return {"config": "C5 64 clips of 1920x1080 -> 860x483, VP pipeline (TP_NUM=60), one GPU", "clips": n_clips,
            "ms_per_frame_step": tot / steps, "clip_frames_per_s": n_clips * steps / (tot * 1e-3),
            "rows_per_clip_mean": float(np.mean([len(x) for x in rows])),
            "note": "the unmodified script runs one clip at ~27-31 frames/s on the CPU (SURVEY 6)"}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4", "c5"]
    res = {}
    for w in which:
        res[w] = {"c3": c3, "c4": c4, "c5": c5}[w]()
        print(json.dumps({w: res[w]}), flush=True)
