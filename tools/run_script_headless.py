"""Run the UNMODIFIED reference scripts headlessly (from /root/reference in the build container, or from the copies
oracle/make_golden.py stages under the git-ignored baseline/_ref/ on the GPU box).

Used by oracle/make_golden.py to produce tests/golden/*.csv -- the reference's own outputs for the
flow -> vanishing-point stage -- by tests/test_gpu_dropin_scripts.py (the scripts under cv.patch(cv2)) and by bench.py's
`dropin` block (wall time of a script with and without the patch).  No arithmetic of the path lives here.

Shims: `imutils.resize` (the module is absent; 8-line INTER_AREA wrapper), a no-op
`matplotlib`, no-op cv2 GUI calls.  The script source is executed as-is via runpy; optional
`overrides` re-bind module-level constants (e.g. TP_NUM) by editing the source TEXT in memory only.
"""
import os
import re
import runpy
import sys
import types


def _install_shims():
    import cv2

    imutils = types.ModuleType("imutils")

    def resize(image, width=None, height=None, inter=cv2.INTER_AREA):
        (h, w) = image.shape[:2]
        if width is None and height is None:
            return image
        if width is None:
            r = height / float(h)
            dim = (int(w * r), height)
        else:
            r = width / float(w)
            dim = (width, int(h * r))
        return cv2.resize(image, dim, interpolation=inter)

    imutils.resize = resize
    sys.modules["imutils"] = imutils

    class _Noop(types.ModuleType):
        """absorbs any attribute access / call / subscript / membership test (matplotlib stand-in)."""

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Noop(name)

        def __call__(self, *a, **k):
            return _Noop("call")

        def __getitem__(self, k):
            return _Noop("item")

        def __contains__(self, k):
            return False

        def __iter__(self):
            return iter(())

        def __bool__(self):
            return False

    mpl = _Noop("matplotlib")
    mpl.pyplot = _Noop("matplotlib.pyplot")
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = mpl.pyplot
    cv2.imshow = lambda *a, **k: None
    # waitKey(10) in the frame loop must not look like ESC/space/backspace; LK_Final.py ends with
    # `while cv.waitKey(0) != 27`, so the blocking form answers ESC
    cv2.waitKey = lambda delay=0, *a, **k: 27 if delay == 0 else 255
    cv2.destroyAllWindows = lambda *a, **k: None


VL_FIELDS = ("m", "intercept", "mv", "interceptv", "bp_x", "bp_y", "lp_y", "rp_y", "up_x", "dp_x")


class VLRecorder:
    """Records the state of the script's `vl` object (class VL, VP.py:196-250; a local of Run()) after every
    `vl.update(...)` call (VP.py:654: once per frame while the VP is initialised) WITHOUT touching the script: a
    sys.setprofile hook looks at the return event of VL.update.  One row per call: VL_FIELDS, NaN while the object is
    not initialised (no VP movement yet)."""

    def __init__(self, script):
        self.script = os.path.abspath(script)
        self.rows = []

    def __call__(self, frame, event, arg):
        if event != "return" or frame.f_code.co_name != "update":
            return
        vl = frame.f_locals.get("self")
        if vl is None or type(vl).__name__ != "VL" or os.path.abspath(frame.f_code.co_filename) != self.script:
            return
        if not vl._isinit:
            self.rows.append([float("nan")] * len(VL_FIELDS))
        else:
            self.rows.append([float(vl._m), float(vl._intercept), float(vl._mv), float(vl._interceptv), float(vl._bp[0]),
                              float(vl._bp[1]), float(vl._lp[1]), float(vl._rp[1]), float(vl._up[0]), float(vl._dp[0])])


def run(script, video_name, workdir, overrides=None, quiet=True, want_vl=False):
    """Execute `script` (path under /root/reference) on `video_name` (a file or %04d pattern inside
    workdir).  Returns the text of ./vps/vps_<name>.csv (and, if want_vl, the VLRecorder rows of the run)."""
    _install_shims()
    os.makedirs(os.path.join(workdir, "vps"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "saved_video"), exist_ok=True)
    cwd, argv = os.getcwd(), sys.argv
    os.chdir(workdir)
    sys.argv = [script, video_name]
    out = sys.stdout
    rec = VLRecorder(script) if want_vl else None
    try:
        if rec is not None:
            sys.setprofile(rec)
        if quiet:
            sys.stdout = open(os.devnull, "w")
        if overrides:
            src = open(script).read()
            for k, v in overrides.items():
                src, n = re.subn(rf"^{k}\s*=.*$", f"{k} = {v!r}", src, count=1, flags=re.M)
                assert n == 1, k
            g = {"__name__": "__main__", "__file__": script}
            exec(compile(src, script, "exec"), g)
        else:
            runpy.run_path(script, run_name="__main__")
    finally:
        sys.setprofile(None)
        sys.stdout = out
        os.chdir(cwd)
        sys.argv = argv
    name = video_name.split("\\")[-1].split(".")[0]
    with open(os.path.join(workdir, "vps", f"vps_{name}.csv"), newline="") as f:
        text = f.read()
    return (text, rec.rows) if want_vl else text
