"""GPU parity of the flow -> vanishing-point stage (a8-a13) and of the whole clip loop, through the C ABI.

  * vp_frame kernel vs oracle/vp_oracle.vp_frame on identical synthetic tracks: CSV rows bit-identical (float64),
    surviving point groups identical -- exercises the checkInside dtype flip (B.1), init aliasing (B.2), the
    +-1 sigma update (B.3), hide/reset (B.4); the vanishing-line pair (class VL) after every frame within 1e-6 px
    (two least-squares fits over <= 300 rows: the CTA's summation order differs from numpy's);
  * ClipBatchRunner (resize -> ... -> CSV) vs the oracle-driven frame loop on a synthetic zoom clip, and against
    the golden CSV written by the UNMODIFIED reference script (tolerance: VP within 1 px, north_star).
"""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _check_vl(state, want, where):
    """state: LkfVpState; want: the oracle's S["vl"] (dict or None).  Slopes to 1e-9 relative, end points to 1e-6 px."""
    from lk_optical_flow_method_b200 import vp as V
    got = V.vl_of_state(state)
    assert (got is None) == (want is None), where
    if want is None:
        return 0
    for k in V.VL_FIELDS:
        assert np.isclose(got[k], want[k], rtol=1e-9, atol=1e-6, equal_nan=True), (where, k, got[k], want[k])
    return 1


def _masks():
    z = np.load(os.path.join(GOLD, "roi_masks_860x483.npz"))
    return z["mask"], list(z["small"])


def _random_tracks(rng, mask, n, foe, lose=0.15, outside=0.1, still=0.05):
    ys, xs = np.nonzero(mask)
    idx = rng.integers(0, len(xs), n)
    p0 = np.stack([xs[idx], ys[idx]], 1).astype(np.float32) + rng.uniform(0, 1, (n, 2)).astype(np.float32)
    d = p0 - np.asarray(foe, np.float32)
    p1 = (p0 + d * np.float32(0.02) * rng.uniform(0.2, 3.0, (n, 1)).astype(np.float32)
          + rng.normal(0, 0.3, (n, 2)).astype(np.float32)).astype(np.float32)
    out = rng.uniform(size=n) < outside
    p1[out] += rng.choice([-1, 1], (out.sum(), 2)).astype(np.float32) * np.float32(200)
    p1[:, 0] = np.clip(p1[:, 0], -20, mask.shape[1] - 1)
    p1[:, 1] = np.clip(p1[:, 1], -20, mask.shape[0] - 1)
    same = rng.uniform(size=n) < still
    p1[same] = p0[same]
    st = (rng.uniform(size=n) >= lose).astype(np.uint8)
    return p0, p1, st


@pytest.mark.parametrize("variant,tp_num,seed,vp_ref", [("VP", 20, 0, 300), ("LKF", 20, 1, 300), ("VP", 100, 2, 300), ("VP", 20, 3, 300),
                                                        ("VP", 20, 4, 0)])
def test_vp_frame_kernel_matches_oracle(variant, tp_num, seed, vp_ref):
    # vp_ref = 0: the reference documents VP_REF = 0 as "use all VPs" (all_vp[-0:] is the whole history)
    import torch
    from lk_optical_flow_method_b200 import vp as V
    from oracle import vp_oracle as O
    mask, small = _masks()
    B, frames = 3, 140
    r = V.ClipBatchRunner(B, (720, 1280), variant=variant,   # Hei = int(860*720/1280) = 483 like the script
                          masks=(mask, small), TP_NUM=tp_num, VP_REF=vp_ref, expected_frames=frames + 8)
    p = O.params(variant, TP_NUM=tp_num, VP_REF=vp_ref)
    S = [O.new_state(p) for _ in range(B)]
    rng = np.random.default_rng(seed)
    maxp = r.maxp
    md = V.image_desc(r.mask)
    n_vl = 0
    for f in range(frames):
        quiet = 60 <= f < 125 and seed == 3                       # long gap: exercises the hide/reset path
        P0 = np.zeros((B, 2, maxp, 2), np.float32)
        P1 = np.zeros_like(P0)
        ST = np.zeros((B, 2, maxp), np.uint8)
        C = np.zeros((B, 2), np.int32)
        host_tracks = []
        for b in range(B):
            tr = []
            for g in range(2):
                n = 0 if (quiet or rng.uniform() < 0.05) else int(rng.integers(1, maxp + 1))
                foe = (412 + 5 * b + rng.normal(0, 2), 290 + rng.normal(0, 2))
                p0, p1, st = _random_tracks(rng, mask, n, foe)
                if n and rng.uniform() < 0.5:
                    st[:] = 1                                     # all tracked: the boolean-mask branch of B.1
                P0[b, g, :n], P1[b, g, :n], ST[b, g, :n], C[b, g] = p0, p1, st, n
                tr.append((p0, p1, st))
            host_tracks.append(tr)
        r.p0.copy_(torch.from_numpy(P0.reshape(B, 2 * maxp, 2)))
        r.p1.copy_(torch.from_numpy(P1.reshape(B, 2 * maxp, 2)))
        r.st.copy_(torch.from_numpy(ST.reshape(B, 2 * maxp)))
        r.counts.copy_(torch.from_numpy(C.reshape(-1)))
        from lk_optical_flow_method_b200._lib import check
        from lk_optical_flow_method_b200.device import stream_ptr
        check(r.ws.lib.lkf_vp_frame(ctypes.byref(r.vp_params), r.states.data_ptr(), B, r.p0.data_ptr(), r.p1.data_ptr(),
                                    r.st.data_ptr(), r.counts.data_ptr(), maxp, ctypes.byref(md), r.rows.data_ptr(), r.row_cap,
                                    r.refresh.data_ptr(), None, stream_ptr()), "lkf_vp_frame")
        new_counts = r.counts.cpu().numpy().reshape(B, 2)
        new_p0 = r.p0.cpu().numpy().reshape(B, 2, maxp, 2)
        refresh = r.refresh.cpu().numpy()
        for b in range(B):
            tracks, lens = [], []
            any_pts = any(len(t[0]) for t in host_tracks[b])
            for g, (p0, p1, st) in enumerate(host_tracks[b]):
                if len(p0) == 0 or not any_pts:
                    tracks.append((np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32)))
                    lens.append(len(p0))
                    continue
                s = O.apply_roi_filter(p1.reshape(-1, 1, 2), st.reshape(-1, 1).copy(), mask)
                keep = s.ravel() == 1
                tracks.append((p0[keep], p1[keep]))
                lens.append(int(keep.sum()))
                assert new_counts[b, g] == keep.sum()
                assert np.array_equal(new_p0[b, g, :keep.sum()], p1[keep])
            O.vp_frame(S[b], tracks if any_pts else [], p, 483)
            want_refresh = (lens[0] + lens[1]) < p["TP_NUM"] * p["TP_UPDATE_RATE"] or S[b]["tp_ult"] == p["TP_UPDATE_TIME"]
            assert bool(refresh[b]) == bool(want_refresh)
            if want_refresh:
                S[b]["tp_ult"] = 0
            S[b]["tp_ult"] += 1
            S[b]["vp_ult"] += 1
        sts = r.state_structs()
        for b in range(B):
            if sts[b].n_rows != len(S[b]["rows"]) and os.path.isdir(os.path.join(ROOT, "gpurun_out")):
                import pickle
                got_rows = r.rows.cpu().numpy()[b, :sts[b].n_rows]
                with open(os.path.join(ROOT, "gpurun_out", "vp_mismatch.pkl"), "wb") as fh:
                    pickle.dump(dict(frame=f, clip=b, tracks=host_tracks[b], gpu_rows=got_rows, ref_rows=S[b]["rows"],
                                     gpu=(sts[b].n_rows, sts[b].n_lines, sts[b].n_cps, sts[b].vp_x, sts[b].vp_y, sts[b].vp_init,
                                          list(sts[b].avg_len)),
                                     ref=(len(S[b]["rows"]), S[b]["n_lines"], S[b]["n_cps"], S[b]["vp"], S[b]["avg"])), fh)
            assert sts[b].n_rows == len(S[b]["rows"]), (f, b, sts[b].n_rows, len(S[b]["rows"]), sts[b].n_lines, S[b]["n_lines"],
                                                         sts[b].n_cps, S[b]["n_cps"], host_tracks[b])
            n_vl += _check_vl(sts[b], S[b]["vl"], (f, b))
            assert bool(sts[b].vp_moved) == bool(S[b]["moved"])
    assert n_vl > 50                                               # the VL path actually ran
    got = r.rows_host()
    states = r.state_structs()
    total = 0
    for b in range(B):
        want = np.array(S[b]["rows"], np.float64).reshape(-1, 2)
        assert got[b].shape == want.shape, (got[b].shape, want.shape)
        assert np.array_equal(got[b], want, equal_nan=True)        # bit-identical CSV rows
        assert states[b].n_lines == S[b]["n_lines"] and states[b].n_cps == S[b]["n_cps"]
        total += len(want)
    assert total > 50
    from lk_optical_flow_method_b200.vp import csv_text
    assert csv_text(got[0]) == O.csv_text(S[0]["rows"])           # a13: same bytes, \r\n line ends
    assert csv_text(got[0]).startswith("x,y\r\n")


def test_resize_area_bit_exact(oracle):
    import torch
    from lk_optical_flow_method_b200 import ops, synth
    tex = synth.make_texture(320, 180, 4)
    f = np.stack([synth.zoom_frame(tex, 320, 180, t) for t in (2, 9)])
    d = torch.from_numpy(f).cuda()
    for (wd, hd) in [(215, 120), (200, 113), (107, 61)]:
        got = ops.resize_area(d, (wd, hd)).cpu().numpy()
        assert np.array_equal(got, np.stack([oracle.resize_area(x, (wd, hd)) for x in f]))
    g = torch.from_numpy(np.ascontiguousarray(f[..., 1])).cuda()
    assert np.array_equal(ops.resize_area(g, (215, 120)).cpu().numpy(), np.stack([oracle.resize_area(x, (215, 120)) for x in f[..., 1]]))
    # the scripts' own geometries (VP.py:430-433): 1280x720 and 1920x1080 frames to WID = 860 (column-strip kernel,
    # word loads; the last columns take its byte path), and a crop whose rows are not 4-byte aligned (per-pixel kernel)
    for (W, H) in [(1280, 720), (1920, 1080)]:
        tex2 = synth.make_texture(W, H, 5)
        fr = np.stack([synth.zoom_frame(tex2, W, H, t) for t in (1, 4)])
        got = ops.resize_area(torch.from_numpy(fr).cuda(), (860, 483)).cpu().numpy()
        assert np.array_equal(got, np.stack([oracle.resize_area(x, (860, 483)) for x in fr])), (W, H)
    crop = torch.from_numpy(fr).cuda()[:, 3:723, 5:1286]                 # 1281 x 720 view, pitch 5760, base offset 15 bytes
    want = np.stack([oracle.resize_area(np.ascontiguousarray(x), (860, 483)) for x in fr[:, 3:723, 5:1286]])
    assert np.array_equal(ops.resize_area(crop, (860, 483)).cpu().numpy(), want)


def _clip(meta, n):
    from lk_optical_flow_method_b200 import synth
    tex = synth.make_texture(meta["W"], meta["H"], meta["clip"].get("seed", 0))
    kw = {k: v for k, v in meta["clip"].items() if k != "seed"}
    return [synth.zoom_frame(tex, meta["W"], meta["H"], t, **kw) for t in range(n)]


@pytest.mark.parametrize("name", ["vp_zoom60", "lkf_zoom60", "vp_zoom60_tp100", "vp_fast40"])
def test_clip_runner_vs_oracle_loop_and_golden(oracle, name):
    import torch
    from lk_optical_flow_method_b200 import vp as V
    from oracle import vp_oracle as O
    meta = json.load(open(os.path.join(GOLD, f"{name}.json")))
    frames = _clip(meta, meta["frames"])
    variant = "LKF" if meta["script"].startswith("LK_Final") else "VP"
    ov = meta["overrides"]
    masks = _masks()
    # two copies of the clip in one batch: clips are independent, results must be identical
    batches = (torch.from_numpy(np.stack([f, f])).cuda() for f in frames)
    rows, runner = V.run_clips(batches, variant=variant, masks=masks, expected_frames=meta["frames"] + 4, **ov)
    assert np.array_equal(rows[0], rows[1])
    ref = O.run_clip(frames, oracle.OracleCV, variant=variant, masks=masks, **ov)
    want = np.array(ref["rows"], np.float64).reshape(-1, 2)
    assert rows[0].shape == want.shape
    assert np.array_equal(rows[0], want)                           # same arithmetic end to end -> same rows
    gold = np.loadtxt(os.path.join(GOLD, f"{name}.csv"), delimiter=",", skiprows=1).reshape(-1, 2)
    assert len(gold) == meta["rows"]
    # against the unmodified script (cv2 arithmetic): VP within 1 px (north_star); rows line up one-to-one
    assert rows[0].shape == gold.shape, (rows[0].shape, gold.shape)
    if len(gold):
        assert np.abs(rows[0] - gold).max() <= 1.0
    # vanishing-line pair after the last frame: against the oracle loop (same rows -> same fits) and against the state
    # of the unmodified script's own `vl` object (tests/golden/vl_*.npz; its rows differ by the cv2-vs-oracle LK
    # difference, so end points within 1 px like the VP itself)
    st = runner.state_structs()
    assert _check_vl(st[0], ref["vl"], name) == (1 if ref["vl"] is not None else 0)
    assert _check_vl(st[1], ref["vl"], name) == (1 if ref["vl"] is not None else 0)
    z = np.load(os.path.join(GOLD, f"vl_{name}.npz"))
    assert len(z["rows"]) == len(ref["vl_rows"]) == meta["vl_calls"]
    if len(z["rows"]) and not np.isnan(z["rows"][-1, 0]):
        got = V.vl_of_state(st[0])
        last = dict(zip([str(k) for k in z["fields"]], z["rows"][-1]))
        assert abs(got["bp_x"] - last["bp_x"]) <= 1.0 and abs(got["bp_y"] - last["bp_y"]) <= 1.0
        lp, rp, up, dp, ok = V.vanishing_lines(got, runner.WID, runner.Hei)
        assert ok and lp[0] == 0 and rp[0] == runner.WID - 1 and up[1] == 0 and dp[1] == runner.Hei - 1
        assert got["lp_y"] == got["bp_y"] - got["bp_x"] * got["m"]


def test_c5_batch_of_1080p_clips_sharded_with_csv(oracle, tmp_path):
    """C5 shape: a batch of independent 1920x1080 clips (per-clip FOE jitter and zoom rate), resized to WID=860 like the
    scripts (VP.py:430-433), processed in lock-step on one GPU, rows gathered per clip and written as vps_<name>.csv.
    Every clip's rows must equal the oracle-driven loop on the same frames (clips are independent: batching and the
    clip -> rank assignment must not change a single row)."""
    import torch
    from lk_optical_flow_method_b200 import shard, synth, vp as V
    from oracle import vp_oracle as O
    W, H, n_clips, n_frames = 1920, 1080, 3, 14
    clips = []
    for c in range(n_clips):
        rng = np.random.default_rng(c)
        foe = (0.48 * (1 + rng.uniform(-0.05, 0.05)), 0.60 * (1 + rng.uniform(-0.05, 0.05)))
        rate = float(rng.uniform(1.010, 1.020))
        tex = synth.make_texture(W, H, c)
        clips.append([synth.zoom_frame(tex, W, H, t, rate=rate, foe=foe) for t in range(n_frames)])
    # the rank's shard: with world_size 1 every clip is local (the 2-rank split is covered by test_shard_gloo)
    mine = shard.assign(n_clips, 1, 0)
    batches = (torch.from_numpy(np.stack([clips[c][t] for c in mine])).cuda() for t in range(n_frames))
    rows, runner = V.run_clips(batches, variant="VP", expected_frames=n_frames + 4, TP_NUM=60)
    assert (runner.Hei, runner.WID) == (483, 860) and runner.resize_needed
    merged = shard.gather_to_root({c: rows[i] for i, c in enumerate(mine)})
    paths = shard.write_vp_csvs(merged, [f"clip{c:02d}" for c in range(n_clips)], str(tmp_path))
    assert len(paths) == n_clips
    n_rows = 0
    for c in range(n_clips):
        ref = O.run_clip(clips[c], oracle.OracleCV, variant="VP", TP_NUM=60)
        want = np.array(ref["rows"], np.float64).reshape(-1, 2)
        assert merged[c].shape == want.shape and np.array_equal(merged[c], want), c
        text = open(paths[c], newline="").read()
        assert text == V.csv_text(want) and text.startswith("x,y\r\n")
        n_rows += len(want)
    assert n_rows > 0                                   # the VP state machine actually produced rows


def test_drop_in_module_in_the_scripts_frame_loop(oracle):
    """The cv2-signature module (lk_optical_flow_method_b200.cv) used the way the scripts use `cv`: numpy frames in,
    numpy results out, four masked goodFeaturesToTrack calls per refresh, one calcOpticalFlowPyrLK per point group,
    `st` mutated in place (VP.py:503, 545-553, 707) -- driven by the restated frame loop.  Same rows as the loop on the
    oracle's functions, within 1 px of the unmodified script's CSV."""
    from lk_optical_flow_method_b200 import cv
    from oracle import vp_oracle as O
    name = "vp_fast40"
    meta = json.load(open(os.path.join(GOLD, f"{name}.json")))
    frames = _clip(meta, meta["frames"])
    masks = _masks()
    got = O.run_clip(frames, cv, variant="VP", masks=masks, **meta["overrides"])
    ref = O.run_clip(frames, oracle.OracleCV, variant="VP", masks=masks, **meta["overrides"])
    a, b = np.array(got["rows"], np.float64).reshape(-1, 2), np.array(ref["rows"], np.float64).reshape(-1, 2)
    assert a.shape == b.shape and np.array_equal(a, b)
    gold = np.loadtxt(os.path.join(GOLD, f"{name}.csv"), delimiter=",", skiprows=1).reshape(-1, 2)
    assert a.shape == gold.shape and np.abs(a - gold).max() <= 1.0
    assert O.csv_text(got["rows"]).startswith("x,y\r\n")
