#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the optical-flow hot path on B200 (BASELINE.json metric, config C2).

Workload (one *step* = B frame pairs of a synthetic 1920x1080 forward-zoom clip, per GPU):
  per pair : process_img(new BGR frame) [a1+a2]  ->  pyramid, 3 levels [a3]  ->  sparse pyramidal LK,
             2000 points, 21x21 window, criteria (EPS|COUNT, 10, 0.03) [a4+a7];
  every 10th pair: goodFeaturesToTrack(2000, q=0.01, d=7, bs=7) on the previous frame replaces that
             pair's points [a5+a6]  (the REPlace refresh of VP.py:700-718).
Arms:
  default          : this repo's CUDA path.  `value` = device-resident throughput (CUDA-graph replay,
                     CUDA events, max over ranks); `e2e` = the same work through the host-buffer API
                     (pinned host frames/points -> H2D -> kernels -> D2H of tracks), copies in the timed
                     region.  Adds `roofline` (dominant kernel), `stages`, `cpu_baseline`, `clocks`.
  --impl reference : the reference's own implementation of the path = OpenCV (cv2) on the host cores,
                     same per-pair work, bounded sample; falls back to the oracle port if cv2 is absent.
Multi-GPU: one process per GPU (torchrun); clips are sharded by rank, no data-path collective;
torch.distributed is used only for the barrier and the max-over-ranks of the timing ("scaling": weak).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, NPTS = 1920, 1080, 2000
LK = dict(winSize=(21, 21), maxLevel=2, criteria=(3, 10, 0.03))
GFTT = dict(maxCorners=NPTS, qualityLevel=0.01, minDistance=7, blockSize=7)
REFRESH = 10
N_ADV = 40          # 2 % adversarial points (outside / border) to exercise the status paths
WORKLOAD = ("C2: 1920x1080 sparse PyrLK, 2000 pts/pair, 3-level pyramid (maxLevel=2), 21x21 window, "
            "criteria (EPS|COUNT,10,0.03); per pair process_img(BGR)+pyramid+LK; GFTT(2000,q=0.01,d=7,bs=7) "
            "refresh on every 10th pair")
METRIC = "frame-pairs/s, 1080p PyrLK 2k pts"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_traffic.json")     # written by tools/ncu_traffic.py from an ncu --set full capture


def source_digest():
    """sha256 over the CUDA sources of the library: a traffic figure is only quoted when it was captured on this build."""
    import hashlib
    d = os.path.join(ROOT, "lk-optical-flow-method_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def captured_traffic():
    """{kernel substring: DRAM bytes per 64-pair step} from the committed capture, or {} when the sources changed since."""
    try:
        t = json.load(open(TRAFFIC_FILE))
    except Exception:
        return {}, None
    if t.get("source_digest") != source_digest():
        return {}, f"{os.path.relpath(TRAFFIC_FILE, ROOT)} is stale (captured on source digest {t.get('source_digest')})"
    return t.get("dram_bytes_per_step", {}), f"{os.path.relpath(TRAFFIC_FILE, ROOT)} ({t.get('how', 'ncu --set full')})"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------
def make_inputs(B, seed, device):
    """B+1 consecutive frames of the zoom clip, generated in HBM; points per pair from the product GFTT
    (parity-tested) + adversarial points."""
    import torch
    from lk_optical_flow_method_b200 import ops, synth
    tex = torch.from_numpy(synth.make_texture(W, H, seed)).to(device)
    frames = torch.stack([synth.zoom_frame(tex, W, H, 1 + t, period=50) for t in range(B + 1)])
    del tex
    gray = ops.process_img(frames[:B])
    pts = torch.zeros((B, NPTS, 2), dtype=torch.float32, device=device)
    adv = torch.from_numpy(synth.adversarial_points(W, H, N_ADV, seed)).to(device)
    for i0 in range(0, B, 8):
        g = gray[i0:i0 + 8]
        eig = ops.min_eigen(g, GFTT["blockSize"])
        cor, cnt = ops.gftt_select(eig, None, NPTS - N_ADV, GFTT["qualityLevel"], GFTT["minDistance"], NPTS - N_ADV)
        cnt = cnt.cpu().numpy()
        for j in range(g.shape[0]):
            k, need = int(cnt[j]), NPTS - N_ADV
            assert k >= need // 2, f"synthetic frame yields only {k} corners"
            pts[i0 + j, :k] = cor[j, :k]
            if k < need:                      # strongly zoomed frames: top up with half-pixel shifted copies
                pts[i0 + j, k:need] = cor[j, :need - k] + 0.5
            pts[i0 + j, need:] = adv
    torch.cuda.synchronize()
    return frames, pts


def bench_config(B):
    """`config` of the JSON line: the same dict in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "pairs_per_step_per_gpu": int(B), "points_per_pair": NPTS,
            "l2": f"inputs larger than L2: {int(B)} BGR frames = {int(B) * 3 * W * H / 1e6:.0f} MB per step"}


def lk_ops(iters_total, n_points, n_levels, win):
    """SURVEY 8d op model: per (point, level) extraction 27 w h; per iteration 12 w h; level-0 error 9 w h."""
    wh = win[0] * win[1]
    return float(n_points) * wh * (27 * n_levels + 9) + float(iters_total) * 12 * wh


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's implementation (cv2) or the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_pairs_per_s(frames_bgr, pts, budget_s, max_pairs=None, threads=None):
    """frames_bgr: list of B+1 HxWx3 numpy frames; pts: (B,N,2).  Runs the per-pair work of the workload
    with cv2 (kind 'reference') or the oracle (kind 'port') for ~budget_s seconds in THIS process with
    `threads` cv2 threads (default: all cores); returns a dict."""
    try:
        import cv2
        kind = "reference"
        cores = threads or os.cpu_count() or 1
        cv2.setNumThreads(cores)
        cores = cv2.getNumThreads()

        def prep(f):
            return cv2.GaussianBlur(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (3, 3), 0)

        def gftt(g):
            return cv2.goodFeaturesToTrack(g, mask=None, **GFTT)

        def lk(a, b, p):
            return cv2.calcOpticalFlowPyrLK(a, b, p, None, **LK)
        impl = f"cv2 {cv2.__version__}"
    except ImportError:
        from oracle import oracle as O
        kind, cores, impl = "port", 1, "oracle/lkflow_oracle.c"

        def prep(f):
            return O.blur3(O.bgr2gray(f))

        def gftt(g):
            return O.good_features_to_track(g, GFTT["maxCorners"], GFTT["qualityLevel"], GFTT["minDistance"],
                                            blockSize=GFTT["blockSize"])

        def lk(a, b, p):
            return O.pyrlk(a, b, p, **LK)
    B = len(frames_bgr) - 1
    prev = prep(frames_bgr[0])
    lk(prev, prep(frames_bgr[1]), pts[0].reshape(-1, 1, 2))      # warm-up (thread pool, caches)
    done, t0 = 0, time.perf_counter()
    while True:
        i = done % B
        if i == 0:
            prev = prep(frames_bgr[0])
        p = pts[i].reshape(-1, 1, 2)
        if i % REFRESH == 0:
            c = gftt(prev)
            if c is not None:
                p = c
        nxt = prep(frames_bgr[i + 1])
        lk(prev, nxt, p)
        prev = nxt
        done += 1
        el = time.perf_counter() - t0
        if el >= budget_s or (max_pairs and done >= max_pairs):
            break
    return {"value": done / el, "unit": "frame-pairs/s", "cores": int(cores), "kind": kind,
            "sample": f"{done} pairs of the same clip/points in {el:.1f} s ({impl}, {cores} thread(s), one process)",
            "pairs": done, "seconds": el}


def cpu_worker(path, seconds):
    """Child process of cpu_pairs_per_s_all_cores: one single-threaded cv2 stream over the shared sample."""
    z = np.load(path, mmap_mode="r")
    frames = [np.ascontiguousarray(f) for f in z["frames"]]
    r = cpu_pairs_per_s(frames, np.ascontiguousarray(z["pts"]), seconds, threads=1)
    print(json.dumps({"pairs": r["pairs"], "seconds": r["seconds"], "kind": r["kind"]}))
    return 0


def cpu_pairs_per_s_all_cores(frames_bgr, pts, budget_s):
    """The reference's CPU path with every host core busy, measured both ways BASELINE.md section 3 names:
    (a) one stream with cv2's own thread pool on all cores, (b) one single-threaded stream per core (independent
    frame pairs, like the GPU arm's batch).  Returns the better of the two as `value`, both in `variants`."""
    import tempfile
    cores = os.cpu_count() or 1
    a = cpu_pairs_per_s(frames_bgr, pts, max(3.0, budget_s / 3))
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, f"lkf_cpu_sample_{os.getpid()}.npz")
    np.savez(path, frames=np.stack(frames_bgr), pts=np.asarray(pts))
    try:
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", path, "--cpu-seconds", str(budget_s)],
                                  stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(cores)]
        outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    finally:
        try:
            os.remove(path)
        except OSError:
            pass
    pairs = sum(o["pairs"] for o in outs)
    secs = max(o["seconds"] for o in outs)
    b = {"value": pairs / secs, "pairs": pairs, "seconds": secs}
    best = b if b["value"] >= a["value"] else a
    return {"value": best["value"], "unit": "frame-pairs/s", "cores": cores, "kind": a["kind"],
            "sample": (f"{pairs} pairs in {secs:.1f} s with {cores} single-threaded cv2 processes (one stream per core); "
                       f"one stream with cv2's {a['cores']}-thread pool: {a['value']:.1f} pairs/s"),
            "variants": {"one_process_all_threads": a["value"], "one_process_per_core": b["value"]},
            "pairs": best["pairs"], "seconds": best["seconds"]}


def run_reference(args):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return 0
    import torch
    from lk_optical_flow_method_b200 import synth
    B = min(args.pairs, 20)
    tex = synth.make_texture(W, H, 0)
    frames = [synth.zoom_frame(tex, W, H, 1 + t, period=50) for t in range(B + 1)]
    try:
        import cv2
        g = [cv2.GaussianBlur(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (3, 3), 0) for f in frames[:B]]
        cor = [cv2.goodFeaturesToTrack(x, mask=None, **dict(GFTT, maxCorners=NPTS - N_ADV)) for x in g]
    except ImportError:
        from oracle import oracle as O
        g = [O.blur3(O.bgr2gray(f)) for f in frames[:B]]
        cor = [O.good_features_to_track(x, NPTS - N_ADV, GFTT["qualityLevel"], GFTT["minDistance"],
                                        blockSize=GFTT["blockSize"]) for x in g]
    adv = synth.adversarial_points(W, H, N_ADV, 0)
    pts = np.stack([np.concatenate([c.reshape(-1, 2), adv])[:NPTS] for c in cor]).astype(np.float32)
    per_step = max(3.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_pairs_per_s(frames, pts, 2.0)
    vals, tot_pairs, tot_s = [], 0, 0.0
    r = None
    for _ in range(args.steps):
        r = cpu_pairs_per_s_all_cores(frames, pts, per_step)
        tot_pairs += r["pairs"]; tot_s += r["seconds"]
    v = tot_pairs / tot_s
    line = {"metric": METRIC, "value": v, "unit": "frame-pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 patches + f32 solves", "data": "synthetic",
            "impl": "reference",
            "config": bench_config(args.pairs),
            "reference_step": {"pairs_per_step": tot_pairs / max(1, args.steps),
                               "note": "each step is a bounded sample (fixed wall budget) of the same per-pair work"},
            "gpus_used": 0,
            "cpu_baseline": {"value": v, "unit": "frame-pairs/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": f"{tot_pairs} pairs in {tot_s:.1f} s over {args.steps} steps; last step: {r['sample']}",
                             "variants": r["variants"]},
            "e2e": {"value": v, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# C5: 64 independent 1080p clips sharded by clip over the ranks, VP CSV per clip
# ------------------------------------------------------------------------------------------------
def c5_clip_frames(c, n_frames, texs, dev):
    import torch
    from lk_optical_flow_method_b200 import synth
    rng = np.random.default_rng(c)
    foe = (0.48 * (1 + rng.uniform(-0.05, 0.05)), 0.60 * (1 + rng.uniform(-0.05, 0.05)))
    rate = float(rng.uniform(1.010, 1.020))
    return [synth.zoom_frame(texs[c % len(texs)], W, H, t, rate=rate, foe=foe) for t in range(n_frames)]


def c5_prepare(clips, texs, n_frames, dev):
    """frames of the clips resident in HBM as per-frame batches + a runner whose step graphs are captured (untimed set-up)"""
    import torch
    from lk_optical_flow_method_b200 import vp as V
    per_clip = [c5_clip_frames(c, n_frames, texs, dev) for c in clips]
    batches = [torch.stack([per_clip[i][t] for i in range(len(clips))]) for t in range(n_frames)]
    del per_clip
    r = V.ClipBatchRunner(len(clips), (H, W), variant="VP", expected_frames=4 * n_frames, TP_NUM=60)
    buf = torch.empty_like(batches[0])
    r.start(batches[0])
    for fr in batches[1:3]:                               # one step per ping-pong parity: both graphs captured
        buf.copy_(fr)
        r.step_graph(buf)
    torch.cuda.synchronize()
    r.reset()
    return r, buf, batches


def c5_run(clips, r, buf, batches):
    """The clips in lock-step on this GPU: ({clip: rows}, seconds from the first frame to the rows on the host)."""
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r.start(batches[0])
    for fr in batches[1:]:
        buf.copy_(fr)
        r.step_graph(buf)
    rows = r.rows_host()                                  # synchronises, D2H of every clip's rows
    dt = time.perf_counter() - t0
    return {c: rows[i] for i, c in enumerate(clips)}, dt


def c5_block(args, rank, world, dev):
    """BASELINE.json config 5: 64 clips of 1920x1080 (-> 860x483 like the scripts), clip c on rank c mod N, every rank runs
    its clips in lock-step; the per-clip rows are gathered to rank 0 over a gloo group (host gather, no NCCL on the data
    path) and written as vps/vps_<clip>.csv.  Timed: first frame -> CSV files on disk, max over ranks.  Rank 0 then repeats
    all 64 clips alone and compares the CSV bytes."""
    import hashlib
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from lk_optical_flow_method_b200 import shard, synth
    n_clips, n_frames = 64, 33
    texs = [torch.from_numpy(synth.make_texture(W, H, 100 + k)).to(dev) for k in range(4)]
    mine = shard.assign(n_clips, world, rank)
    host_group = dist.new_group(backend="gloo") if world > 1 else None
    runner, buf, batches = c5_prepare(mine, texs, n_frames, dev)
    out_dir = tempfile.mkdtemp(prefix="lkf_c5_") if rank == 0 else None
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    rows, dev_s = c5_run(mine, runner, buf, batches)
    n_rows_local = sum(len(v) for v in rows.values())
    texts = shard.render_csvs(rows)                       # every rank formats its own clips
    merged = shard.gather_to_root(texts, group=host_group) if world > 1 else texts
    sha = None
    if rank == 0:
        paths = shard.write_vp_csvs(merged, [f"clip{c:02d}" for c in range(n_clips)], out_dir)
        h = hashlib.sha256()
        for pth in paths:
            h.update(open(pth, "rb").read())
        sha = h.hexdigest()
    if world > 1:
        dist.barrier()
    total_s = time.perf_counter() - t0
    t = torch.tensor([total_s, dev_s], dtype=torch.float64, device=dev)
    nr = torch.tensor([n_rows_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nr, op=dist.ReduceOp.SUM)
    res = None
    if rank == 0:
        total_s, dev_s = float(t[0]), float(t[1])
        n_rows = int(nr.item())
        res = {"clips": n_clips, "frames_per_clip": n_frames, "ranks": world, "clips_per_rank": len(mine),
               "clip_frames_per_s": n_clips * n_frames / total_s, "seconds": total_s, "device_seconds_max_rank": dev_s,
               "csv_rows": n_rows, "csv_sha256": sha, "gather": "every rank renders its clips' CSV text; gloo gather_object to rank 0, which writes one file per clip (save_csv schema)",
               "timed": "first frame of every clip -> CSV files written; frames resident in HBM, step graphs captured beforehand; max "
                        "over ranks (wall clock: it includes the D2H of the rows, the host gather and the file writes)"}
        if world > 1:
            del runner, buf, batches
            allc = list(range(n_clips))
            ref_rows, _ = c5_run(allc, *c5_prepare(allc, texs, n_frames, dev))
            d2 = tempfile.mkdtemp(prefix="lkf_c5ref_")
            h = hashlib.sha256()
            for pth in shard.write_vp_csvs(ref_rows, [f"clip{c:02d}" for c in range(n_clips)], d2):
                h.update(open(pth, "rb").read())
            shutil.rmtree(d2, ignore_errors=True)
            res["csv_equal_to_1gpu_run"] = h.hexdigest() == sha
        shutil.rmtree(out_dir, ignore_errors=True)
    if world > 1:
        dist.barrier()
    return res


# ------------------------------------------------------------------------------------------------
# drop-in boundary: per-call latency of the cv2-signature module with HOST arrays, cv2 beside it
# ------------------------------------------------------------------------------------------------
def dropin_block(args):
    import cv2
    import torch
    from lk_optical_flow_method_b200 import cv as b200, synth, vp as V
    out = {}

    def script_loop(mod, frames, masks, fp, lk, n_groups=2):
        """the scripts' frame loop reduced to its four hot calls (VP.py:404-425, 500-507, 545-553, 682, 700-711)"""
        tm = {"process_img": [0.0, 0], "goodFeaturesToTrack": [0.0, 0], "calcOpticalFlowPyrLK": [0.0, 0]}

        def timed(key, fn):
            t0 = time.perf_counter()
            r = fn()
            tm[key][0] += time.perf_counter() - t0
            tm[key][1] += 1
            return r

        def detect(img):
            groups = []
            for j in range(2):
                p0 = []
                for i in range(2):
                    c = timed("goodFeaturesToTrack", lambda: mod.goodFeaturesToTrack(img, mask=masks[j * 2 + i], **fp))
                    if c is not None:
                        p0.extend(c)
                groups.append(np.array(p0, dtype=np.float32).reshape(-1, 1, 2))
            return groups

        old = timed("process_img", lambda: mod.GaussianBlur(mod.cvtColor(frames[0], cv2.COLOR_BGR2GRAY), (3, 3), 0))
        groups = detect(old)
        t_all = time.perf_counter()
        for k, f in enumerate(frames[1:], 1):
            new = timed("process_img", lambda: mod.GaussianBlur(mod.cvtColor(f, cv2.COLOR_BGR2GRAY), (3, 3), 0))
            nxt = []
            for g in groups:
                if len(g) == 0:
                    nxt.append(g)
                    continue
                p1, st, err = timed("calcOpticalFlowPyrLK", lambda: mod.calcOpticalFlowPyrLK(old, new, g, None, **lk))
                nxt.append(np.ascontiguousarray(p1[st.ravel() == 1].reshape(-1, 1, 2)))
            groups = nxt
            old = new.copy()
            if k % 10 == 0:
                groups = detect(old)
        wall = time.perf_counter() - t_all
        res = {k: 1e6 * v[0] / max(1, v[1]) for k, v in tm.items()}
        res["frame_ms"] = 1e3 * wall / (len(frames) - 1)
        return res

    cv2.setNumThreads(os.cpu_count() or 1)
    for tag, (w, h, n, fp, lk) in {
            "script_860x483": (860, 483, 41, dict(maxCorners=5, qualityLevel=0.3, minDistance=7, blockSize=7),
                               dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03))),
            "c2_1920x1080": (W, H, 21, dict(maxCorners=500, qualityLevel=0.01, minDistance=7, blockSize=7), LK)}.items():
        tex = synth.make_texture(w, h, 11)
        frames = [synth.zoom_frame(tex, w, h, t) for t in range(n)]
        _, masks = V.roi_masks(w, h)
        r = {"us_per_call": {}, "points": f"{fp['maxCorners']} corners per ROI mask, 2 point groups"}
        r["us_per_call"]["cv2"] = script_loop(cv2, frames, masks, fp, lk)
        b200.cache_enable(True)
        script_loop(b200, frames[:3], masks, fp, lk)                 # warm-up (workspace, kernel loading)
        b200.cache_enable(False)
        r["us_per_call"]["b200_uncached"] = script_loop(b200, frames, masks, fp, lk)
        b200.cache_enable(True)
        r["us_per_call"]["b200_cached"] = script_loop(b200, frames, masks, fp, lk)
        r["cache"] = b200.cache_stats()
        out[tag] = r
    # wall time of the unmodified script with and without patch(cv2) (staged copy under baseline/_ref, see tests)
    script = os.path.join(ROOT, "baseline", "_ref", "VP_detection_using_optical_flow.py")
    if os.path.exists(script):
        import shutil
        import tempfile
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import run_script_headless as R
        d = tempfile.mkdtemp(prefix="lkf_dropin_")
        tex = synth.make_texture(1280, 720, 0)
        for t in range(60):
            cv2.imwrite(os.path.join(d, f"f_{t:04d}.png"), synth.zoom_frame(tex, 1280, 720, t))
        res = {}
        for arm in ("cv2", "patched", "patched"):
            saved = b200.patch(cv2) if arm == "patched" else None
            b200.cache_clear()
            t0 = time.perf_counter()
            try:
                text = R.run(script, "f_%04d.png", d)
            finally:
                if saved:
                    b200.unpatch(cv2, saved)
            res[arm] = {"seconds": time.perf_counter() - t0, "csv_rows": text.count("\n") - 1}
        shutil.rmtree(d, ignore_errors=True)
        out["unmodified_VP_script_60_frames_1280x720"] = res
        out["note"] = ("the script's wall time is dominated by PNG decode, imutils.resize and Python glue (out of the path, SURVEY 2); "
                       "per-call numbers are the path")
    return out


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Pin this rank to the CPUs next to its GPU (PCIe root / NUMA node) before the pinned host buffers are allocated, so
    the H2D stream of the e2e arm does not cross sockets.  Best effort: silently does nothing where sysfs says nothing."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        addr = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{addr}: cpus {spec}"
    except Exception:
        pass
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from lk_optical_flow_method_b200 import _lib, ops
    from lk_optical_flow_method_b200.device import Workspace
    from lk_optical_flow_method_b200.pipeline import HostFeeder, PairPipeline
    B, K, Wm = args.pairs, args.steps, max(3, args.warmup)
    ws = Workspace.current()
    frames, pts = make_inputs(B, seed=rank, device=dev)
    pipe = PairPipeline(B, H, W, NPTS, LK, GFTT, REFRESH)
    pipe.prime(frames[0])
    pipe.set_points(pts)
    new_frames = frames[1:]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (eager; also counts kernel launches per step) -----------------------------------
    l0 = ws.launches()
    for i in range(Wm):
        pipe.set_points(pts)
        pipe.step(new_frames)
        if i == 0:
            torch.cuda.synchronize()
            launches_per_step = ws.launches() - l0
    torch.cuda.synchronize()
    replay = pipe.capture(new_frames)
    replay()
    barrier()

    # ---- timed: K graph replays, device-resident ------------------------------------------------
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    smi_index = local
    if vis:
        ent = vis.split(",")
        if local < len(ent) and ent[local].strip().isdigit():
            smi_index = int(ent[local])
    clocks = Clocks(smi_index)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        replay()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * K / (ms_total * 1e-3)

    # ---- sustained leg: >= args.sustain_seconds of back-to-back replays (the K-step region above is a burst of ~20 ms) ---
    n_sus = max(K, int(np.ceil(args.sustain_seconds * 1e3 / (ms_total / K))))
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(n_sus):
        replay()
    s1.record()
    barrier()
    t = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sus_ms = float(t.item())
    sustained = {"value": world * B * n_sus / (sus_ms * 1e-3), "unit": "frame-pairs/s", "steps": int(n_sus),
                 "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus,
                 "note": "same graph replayed back to back for seconds; clocks of the whole run in `clocks`"}

    # ---- e2e: host buffers, copies inside the timed region ---------------------------------------
    host_frames = new_frames.cpu().pin_memory()
    host_pts = pts.cpu().pin_memory()
    feeder = HostFeeder(pipe)
    for _ in range(2):
        feeder.run([(host_frames, host_pts)] * 2)
    barrier()
    t0 = time.perf_counter()
    outs = feeder.run([(host_frames, host_pts)] * K)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / float(t.item())
    h2d = host_frames.numel() + host_pts.numel() * 4
    d2h = sum(o.numel() * o.element_size() for o in outs[-1])
    # ---- plain pinned host -> device copy rate at this N (all ranks at once): the ceiling of the e2e figure ----------
    dst = torch.empty_like(new_frames)
    dst.copy_(host_frames, non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(4):
        dst.copy_(host_frames, non_blocking=True)
    c1.record()
    barrier()
    t = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gbs = world * 4 * host_frames.numel() / (float(t.item()) * 1e-3) / 1e9
    h2d_bw = {"GBps_all_ranks": gbs, "GBps_per_rank": gbs / world, "bytes": int(host_frames.numel()),
              "e2e_h2d_GBps": e2e_value * (h2d / B) / 1e9,
              "note": "cudaMemcpyAsync of one step's pinned BGR frames, 4 in a row, every rank at the same time"}
    del dst
    clk = clocks.stop() if rank == 0 else None
    c5 = c5_block(args, rank, world, dev) if not args.no_c5 else None
    dropin = dropin_block(args) if (rank == 0 and world == 1 and not args.no_dropin) else None

    # ---- per-stage device times: each stage captured in its own CUDA graph and replayed back to back (no launch
    #      gaps inside the timed region), CUDA events on the launch stream --------------------------------------
    stage_names = ["process_img", "pyramid", "gftt_refresh", "lk_track", "carry"]
    stage_fn = [lambda: pipe.stage_process(new_frames), pipe.stage_pyramid, pipe.stage_refresh,
                lambda: pipe.stage_track(want_iters=False), pipe.stage_carry]
    n_step_stages = len(stage_fn)
    pipe.set_points(pts)
    pipe.step(new_frames)                      # state every stage graph replays on: points after one refresh
    pipe.stage_track(want_iters=True)          # iteration counts for the op model (same points as the timed replays)
    torch.cuda.synchronize()
    iters_total = int(pipe.iters.sum().item())
    n_tracked = int(pipe.counts.clamp(min=0).sum().item())
    reps = max(5, min(K, 20))
    st_ms = np.zeros(len(stage_fn))
    for j, fn in enumerate(stage_fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(2):
            g.replay()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b_.record()
        torch.cuda.synchronize()
        st_ms[j] = a.elapsed_time(b_) / reps
        del g

    # ---- FP32 pipe peak probe ------------------------------------------------------------------
    import ctypes
    scratch = torch.zeros(4, device=dev)
    blocks, it = 148 * 16, 2000
    from lk_optical_flow_method_b200.device import stream_ptr
    for _ in range(2):
        ws.lib.lkf_probe_fp32_fma(blocks, it, scratch.data_ptr(), stream_ptr())
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    ws.lib.lkf_probe_fp32_fma(blocks, it, scratch.data_ptr(), stream_ptr())
    p1.record()
    torch.cuda.synchronize()
    fp32_peak = blocks * 256 * it * 128 * 2 / (p0.elapsed_time(p1) * 1e-3) / 1e12

    # ---- rooflines -------------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    P = W * H
    nl = pipe.n_levels
    pyr_bytes = sum((P / 4 ** l) + (P / 4 ** (l + 1)) for l in range(nl - 1))
    R = pipe.R
    alg = {"process_img": B * 4.0 * P, "pyramid": B * pyr_bytes,
           "gftt_refresh": R * (5.0 * P + 9.0 * P), "carry": 2.0 * sum(P / 4 ** l for l in range(nl))}
    stages = {}
    step_sum = st_ms[:n_step_stages].sum()
    for j_, (n_, m_) in enumerate(zip(stage_names, st_ms)):
        d = {"ms": float(m_), "share": float(m_ / step_sum) if j_ < n_step_stages else None}
        if n_ in alg:
            d.update({"bound": "hbm", "algorithmic_bytes": alg[n_], "achieved_GBps": alg[n_] / (m_ * 1e-3) / 1e9,
                      "frac_of_hbm_peak": alg[n_] / (m_ * 1e-3) / 1e9 / hbm_peak})
        stages[n_] = d
    ops_lk = lk_ops(iters_total, n_tracked, nl, LK["winSize"])
    lk_ms = st_ms[stage_names.index("lk_track")]
    lk_tflops = ops_lk / (lk_ms * 1e-3) / 1e12
    stages["lk_track"].update({"bound": "fp32", "algorithmic_ops": ops_lk, "iterations_per_point": iters_total / max(1, n_tracked),
                               "achieved_TFLOPs": lk_tflops, "frac_of_fp32_peak": lk_tflops / fp32_peak})
    traffic, traffic_src = captured_traffic()

    def step_traffic(key):
        v = traffic.get(key)
        return None if v is None else float(v) * B / 64.0

    sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
    fp32_nominal = torch.cuda.get_device_properties(dev).multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    dom = stage_names[int(np.argmax(st_ms[:n_step_stages]))]
    if dom == "lk_track":
        roof = {"kernel": "k_lk_track_t<21,21>", "bound": "fp32", "achieved": lk_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": lk_tflops / fp32_peak, "peak_nominal": fp32_nominal, "frac_of_nominal": lk_tflops / fp32_nominal,
                "traffic": step_traffic("k_lk_track_t"), "traffic_source": traffic_src,
                "peak_source": "measured in this run (FFMA probe: 16 chains x 128 FFMA per loop trip, non-tensor FP32 pipe); "
                               "peak_nominal = SMs x 128 lanes x 2 x the SM clock sampled under load",
                "note": "integer gather/iterate kernel (IMAD/IDP.2A and integer-ALU datapaths run at half the FFMA rate: "
                        "tools/probe_pipes.py); ops from the SURVEY 8d model with on-device iteration counts"}
    else:
        s_ = stages[dom]
        roof = {"kernel": dom, "bound": "hbm", "achieved": s_["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": s_["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src}
    roof_hbm = {"kernel": "k_blur3_stream<RowBGR> (process_img)", "bound": "hbm", "achieved": stages["process_img"]["achieved_GBps"],
                "peak": hbm_peak, "unit": "GB/s", "frac": stages["process_img"]["frac_of_hbm_peak"],
                "traffic": step_traffic("k_blur3_stream"), "traffic_source": traffic_src, "peak_source": peak_src}

    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        # BASELINE.json configs 3 and 4 (parity-test cases, not the bench line): device times with cv2 beside them
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        try:
            import bench_configs
            configs = {"c3": bench_configs.c3(), "c4": bench_configs.c4(),
                       "note": "device-resident inputs, CUDA events, back-to-back repetitions (tools/bench_configs.py); C4 = dense "
                               "per-pixel LK: `gpu_ms` through lkf_pyrlk_sparse with the raster grid as prevPts (switches to the tiled "
                               "dense kernel on the device), `gpu_ms_thread_per_feature_kernel` = the same points in reversed order"}
        except Exception as e:      # the headline line must not depend on the side blocks
            configs = {"error": repr(e)}
    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            hf = [f.numpy() for f in frames.cpu()]
            cpu = cpu_pairs_per_s_all_cores(hf[:21], pts.cpu().numpy()[:20], args.cpu_seconds)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "variants")}
        line = {"metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/int32 patches + f32 solves", "data": "synthetic",
                "config": bench_config(B),
                "timed": "CUDA-graph replay of the whole step, CUDA events on the launch stream; torch.distributed (NCCL) only for "
                         "the barrier and the MAX of the timing, never on the data path",
                "sustained": sustained, "h2d_pinned": h2d_bw, "dropin": dropin, "c5": c5, "configs": configs,
                "e2e": {"value": e2e_value, "unit": "frame-pairs/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h),
                        "note": "pinned host BGR frames + points -> H2D -> kernels -> D2H tracks, double-buffered copy stream",
                        "cpu_binding_rank0": numa},
                "gpu_launches": int(launches_per_step * K), "launches_per_step": int(launches_per_step),
                "roofline": roof, "roofline_hbm_stencil": roof_hbm, "stages": stages,
                "fp32_peak_tflops_measured": fp32_peak, "fp32_peak_tflops_nominal": fp32_nominal, "cpu_baseline": cpu, "clocks": clk}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def _quiet_stdout():
    """stdout carries exactly ONE JSON line: anything a library prints to fd 1 on the way (NCCL's version banner at
    communicator creation, torch.distributed.run notices) is sent to stderr; emit() writes to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="frame pairs per step per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--sustain-seconds", type=float, default=3.0, help="length of the sustained leg (graph replays)")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 block (64 clips sharded over the ranks, CSV output)")
    ap.add_argument("--no-dropin", action="store_true", help="skip the per-call drop-in latencies (N=1 only)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 timing block (N=1 only)")
    ap.add_argument("--cpu-worker", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_worker:
        return cpu_worker(args.cpu_worker, args.cpu_seconds)
    if args.impl == "b200" and args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` on its own: start the N ranks the way the driver does
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "b200" and env_int("WORLD_SIZE", 1) != args.gpus:
        sys.stderr.write(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={env_int('WORLD_SIZE', 1)}; reporting the ranks that exist\n")
    _quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
