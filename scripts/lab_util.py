"""Lab helpers shared by the A/B scripts (never imported by the package, the tests or bench.py)."""
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
PKG = os.path.join(ROOT, "cs3602-llm-inference-acceleration_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def use_lab_library_if_asked():
    """KVC_LAB_LIBRARY=1: point the binding at csrc/libkvc_sm100a_lab.so (scripts/build_lab.sh) before it loads."""
    from kvcompress import _engine

    if os.environ.get("KVC_AB_LIBRARY"):      # any other build of the library, by file name inside csrc/
        _engine._LIB_PATH = os.path.join(PKG, "csrc", os.environ["KVC_AB_LIBRARY"])
    elif os.environ.get("KVC_LAB_LIBRARY"):
        lab = os.path.join(PKG, "csrc", "libkvc_sm100a_lab.so")
        if not os.path.exists(lab):
            raise SystemExit("lab library missing: run scripts/build_lab.sh")
        _engine._LIB_PATH = lab
    return _engine
