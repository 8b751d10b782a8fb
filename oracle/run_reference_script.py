"""Headless runner of the UNMODIFIED reference scripts (shims for imutils / matplotlib / GUI calls).  The implementation
lives in tools/run_script_headless.py (it contains no arithmetic of the path, so bench.py's `dropin` block may use it too);
this module keeps the name oracle/make_golden.py and the tests import."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from run_script_headless import VL_FIELDS, VLRecorder, run, _install_shims  # noqa: E402,F401
