"""Device-resident frame-pair pipeline (the throughput path of BASELINE.json's metric).

One *step* processes B frame pairs of a clip held in HBM: for every new BGR frame
    process_img (a1+a2 fused) -> pyramid (a3) ,
then for every pair (frame i-1 -> frame i)
    pyramidal LK (a4 on the fly + a7)  on that pair's points,
and, on every `refresh_every`-th pair, goodFeaturesToTrack (a5+a6) on the pair's previous frame whose
corners replace that pair's points -- the REPlace refresh of VP.py:700-718 (TP_UPDATE_TIME = 10).
Frame i's pyramid serves as `next` of pair i and `prev` of pair i+1, so each frame is processed once
(the scripts' cv2 calls rebuild both pyramids on every call, VP.py:545).

All launches are enqueued without host synchronisation; `capture()` records the step into a CUDA graph.
Inside a step the corner refresh (a latency-bound chain: one sort/selection CTA per image) runs on a second,
high-priority stream together with the LK launch of the refreshed pairs, concurrently with the LK launch of all
other pairs on the main stream (fork after the pyramids, join before the carry).  The GFTT scratch of the workspace
is only ever touched from that second stream.
"""
import torch

from . import ops
from .device import Workspace, pyramid_shapes


def _in_ws(fn):
    """run a stage method inside the pipeline's own workspace"""
    def wrapped(self, *a, **k):
        with self.ws.use():
            return fn(self, *a, **k)
    wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
    return wrapped


class PairPipeline:
    def __init__(self, B, H, W, n_pts, lk_params=None, gftt_params=None, refresh_every=10, device=None):
        self.B, self.H, self.W, self.N = int(B), int(H), int(W), int(n_pts)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lk = dict(winSize=(21, 21), maxLevel=2, criteria=(3, 10, 0.03), flags=0, minEigThreshold=1e-4)
        if lk_params:
            self.lk.update(lk_params)
        self.gftt = dict(maxCorners=self.N, qualityLevel=0.01, minDistance=7, blockSize=7)
        if gftt_params:
            self.gftt.update(gftt_params)
        self.refresh_every = int(refresh_every)
        self.n_levels = ops.effective_levels(H, W, self.lk["winSize"], self.lk["maxLevel"])
        self.params = ops.lk_params_struct(**self.lk)
        self.ws = Workspace.private()      # own ctx: the captured step holds this workspace's scratch addresses
        # pyramid slots 0..B : slot 0 = last frame of the previous step
        self.levels = [torch.zeros((self.B + 1, h, w), dtype=torch.uint8, device=self.dev)
                       for (h, w) in pyramid_shapes(H, W, self.n_levels)]
        self.pts = torch.zeros((self.B, self.N, 2), dtype=torch.float32, device=self.dev)
        self.counts = torch.full((self.B,), self.N, dtype=torch.int32, device=self.dev)
        self.next_pts = torch.zeros_like(self.pts)
        self.status = torch.zeros((self.B, self.N), dtype=torch.uint8, device=self.dev)
        self.err = torch.zeros((self.B, self.N), dtype=torch.float32, device=self.dev)
        self.iters = None
        self.R = len(range(0, self.B, self.refresh_every)) if self.refresh_every > 0 else 0
        if self.R:
            self.corners = torch.zeros((self.R, self.N, 2), dtype=torch.float32, device=self.dev)
            self.ccounts = torch.zeros((self.R,), dtype=torch.int32, device=self.dev)
            # refreshed pairs are tracked by their own launch (second stream); the main launch skips them
            self.main_mask = torch.ones((self.B,), dtype=torch.int32, device=self.dev)
            self.main_mask[0:self.B:self.refresh_every] = 0
            self.counts_main = torch.zeros_like(self.counts)
            self.r_next = torch.zeros((self.R, self.N, 2), dtype=torch.float32, device=self.dev)
            self.r_status = torch.zeros((self.R, self.N), dtype=torch.uint8, device=self.dev)
            self.r_err = torch.zeros((self.R, self.N), dtype=torch.float32, device=self.dev)
            self.aux = torch.cuda.Stream(device=self.dev, priority=-1)
        self.overlap = True
        self.graph = None

    # -- state -----------------------------------------------------------------------------------
    @_in_ws
    def prime(self, first_bgr):
        """Process the frame preceding the first pair into slot 0 ([H,W,3] uint8 CUDA tensor)."""
        ops.process_img(first_bgr[None] if first_bgr.dim() == 3 else first_bgr, out=self.levels[0][0:1])
        for l in range(1, self.n_levels):
            ops.pyr_down(self.levels[l - 1][0:1], out=self.levels[l][0:1])

    def set_points(self, pts, counts=None):
        self.pts.copy_(pts)
        if counts is not None:
            self.counts.copy_(counts)

    # -- stages (each is one kernel launch or a fixed short sequence) -----------------------------
    @_in_ws
    def stage_process(self, bgr, lo=0, hi=None):
        """process_img of new frames lo..hi-1 (pyramid slots lo+1..hi)."""
        hi = self.B if hi is None else hi
        ops.process_img(bgr[lo:hi], out=self.levels[0][1 + lo:1 + hi])

    @_in_ws
    def stage_pyramid(self, lo=0, hi=None):
        hi = self.B if hi is None else hi
        for l in range(1, self.n_levels):
            ops.pyr_down(self.levels[l - 1][1 + lo:1 + hi], out=self.levels[l][1 + lo:1 + hi])

    @_in_ws
    def stage_refresh(self):
        if not self.R:
            return
        e = self.refresh_every
        # a5+a6 in one ABI call: the masked maximum is reduced while the eig map is produced (map in the workspace)
        ops.gftt_detect(self.levels[0][0:self.B:e], None, self.gftt["maxCorners"], self.gftt["qualityLevel"],
                        self.gftt["minDistance"], self.gftt["blockSize"], self.N, out=(self.corners, self.ccounts))
        self.pts[0:self.B:e].copy_(self.corners)
        self.counts[0:self.B:e].copy_(self.ccounts)

    @_in_ws
    def stage_track(self, want_iters=False):
        prev = [lv[0:self.B] for lv in self.levels]
        nxt = [lv[1:self.B + 1] for lv in self.levels]
        out = ops.pyrlk(prev, nxt, self.pts, self.params, counts=self.counts, want_iters=want_iters,
                        out=(self.next_pts, self.status, self.err))
        if want_iters:
            self.iters = out[3]

    def stage_carry(self):
        for lv in self.levels:
            lv[0].copy_(lv[self.B])

    @_in_ws
    def stage_process_pyramid(self, bgr):
        """process_img + the whole pyramid of the B new frames in one ABI call."""
        ops.process_pyramid(bgr, [lv[1:] for lv in self.levels])

    @_in_ws
    def stage_track_main(self, lo=0, hi=None):
        """LK of pairs lo..hi-1 that are NOT refreshed in this step (point count 0 for the refreshed ones)."""
        hi = self.B if hi is None else hi
        torch.mul(self.counts[lo:hi], self.main_mask[lo:hi], out=self.counts_main[lo:hi])
        prev = [lv[lo:hi] for lv in self.levels]
        nxt = [lv[lo + 1:hi + 1] for lv in self.levels]
        ops.pyrlk(prev, nxt, self.pts[lo:hi], self.params, counts=self.counts_main[lo:hi],
                  out=(self.next_pts[lo:hi], self.status[lo:hi], self.err[lo:hi]))

    @_in_ws
    def stage_refresh_detect(self):
        """Corner refresh of every refresh_every-th pair (needs level 0 of the previous frames only)."""
        e = self.refresh_every
        ops.gftt_detect(self.levels[0][0:self.B:e], None, self.gftt["maxCorners"], self.gftt["qualityLevel"],
                        self.gftt["minDistance"], self.gftt["blockSize"], self.N, out=(self.corners, self.ccounts))

    @_in_ws
    def stage_track_refreshed(self):
        """LK launch of exactly the refreshed pairs (strided views of the pyramid slots), results scattered into the
        per-pair outputs."""
        e = self.refresh_every
        prev = [lv[0:self.B:e] for lv in self.levels]
        nxt = [lv[1:self.B + 1:e] for lv in self.levels]
        ops.pyrlk(prev, nxt, self.corners, self.params, counts=self.ccounts, out=(self.r_next, self.r_status, self.r_err))
        self.pts[0:self.B:e].copy_(self.corners)
        self.counts[0:self.B:e].copy_(self.ccounts)
        self.next_pts[0:self.B:e].copy_(self.r_next)
        self.status[0:self.B:e].copy_(self.r_status)
        self.err[0:self.B:e].copy_(self.r_err)

    def step(self, bgr):
        """bgr: [B,H,W,3] uint8 CUDA tensor = the B new frames of this step."""
        with self.ws.use():
            self._step(bgr)

    def _step(self, bgr):
        if not (self.R and self.overlap):
            self.stage_process_pyramid(bgr)
            self.stage_refresh()
            self.stage_track()
            self.stage_carry()
            return
        # Fork after the stencil pass: the refresh chain and the LK launch of the refreshed pairs run on the
        # second stream next to the main LK launch (the one-CTA-per-image sort/selection is latency-bound).  (Splitting the batch in halves to overlap the stencils of one half
        # with the tracker of the other was measured slower: the stencil CTAs queue behind the tracker's.)
        main = torch.cuda.current_stream()
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        self.stage_process_pyramid(bgr)
        fork.record(main)
        with torch.cuda.stream(self.aux):
            self.aux.wait_event(fork)
            self.stage_refresh_detect()
            self.stage_track_refreshed()
            join.record(self.aux)
        self.stage_track_main()
        main.wait_event(join)
        self.stage_carry()

    # -- CUDA graph ------------------------------------------------------------------------------
    def capture(self, bgr):
        """Record one step on `bgr` (a persistent input buffer) into a CUDA graph; returns replay()."""
        if self.R:
            self.stage_refresh_detect()      # eager: sizes this workspace's scratch (a captured call may not grow it); its
                                             # outputs are rewritten by every replay
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step(bgr)
        self.graph = g
        return g.replay

    def results(self):
        return self.next_pts, self.status, self.err, self.counts


class HostFeeder:
    """Host-buffer front end of PairPipeline: pinned host frames/points in, pinned host tracks out.

    H2D copies run on a dedicated copy stream into one of two device input slots while the previous
    step computes (double buffering); results return through pinned buffers (also two slots, so the
    tuple returned for step k is valid until step k+2 is issued).  The D2H copies run on a third stream: the
    compute stream only snapshots the step's results into a device-side slot (a 1.7 MB device copy), so the next
    step's kernels never wait behind a PCIe transfer."""

    def __init__(self, pipe):
        self.pipe = pipe
        p = pipe
        self.copy_stream = torch.cuda.Stream(device=p.dev)
        self.d2h_stream = torch.cuda.Stream(device=p.dev)
        self.stage = [tuple(torch.empty_like(t) for t in p.results()) for _ in range(2)]     # device-side snapshots of the results
        self.in_frames = [torch.empty((p.B, p.H, p.W, 3), dtype=torch.uint8, device=p.dev) for _ in range(2)]
        self.in_pts = [torch.empty((p.B, p.N, 2), dtype=torch.float32, device=p.dev) for _ in range(2)]
        self.out = [(torch.empty((p.B, p.N, 2), dtype=torch.float32).pin_memory(),
                     torch.empty((p.B, p.N), dtype=torch.uint8).pin_memory(),
                     torch.empty((p.B, p.N), dtype=torch.float32).pin_memory(),
                     torch.empty((p.B,), dtype=torch.int32).pin_memory()) for _ in range(2)]

    def run(self, batches):
        """batches: iterable of (frames [B,H,W,3] uint8, pts [B,N,2] float32) host tensors (pinned for
        asynchronous copies).  Returns the list of per-step (next_pts, status, err, counts) host tuples."""
        p = self.pipe
        compute = torch.cuda.current_stream()
        h2d = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        d2h = [torch.cuda.Event() for _ in range(2)]
        snap = [torch.cuda.Event() for _ in range(2)]
        outs = []
        for k, (frames, pts) in enumerate(batches):
            s = k & 1
            with torch.cuda.stream(self.copy_stream):
                if k >= 2:
                    self.copy_stream.wait_event(done[s])          # slot s is free once step k-2 computed
                self.in_frames[s].copy_(frames, non_blocking=True)
                self.in_pts[s].copy_(pts, non_blocking=True)
                h2d[s].record(self.copy_stream)
            compute.wait_event(h2d[s])
            p.set_points(self.in_pts[s])
            p.counts.fill_(p.N)
            p.step(self.in_frames[s])
            done[s].record(compute)
            if k >= 2:
                d2h[s].synchronize()                              # host slot s consumed (step k-2 landed): its device snapshot is free too
            for dst, src in zip(self.stage[s], p.results()):
                dst.copy_(src)
            snap[s].record(compute)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(snap[s])
                for dst, src in zip(self.out[s], self.stage[s]):
                    dst.copy_(src, non_blocking=True)
                d2h[s].record(self.d2h_stream)
            outs.append(self.out[s])
        for e in d2h:
            e.synchronize()
        return outs
