"""TEST / BENCH INFRASTRUCTURE ONLY — loads the *unmodified* reference (MVIP-NeRF DS_NeRF).

Looks for the tree at $MVIP_REFERENCE_ROOT, /root/reference (build container) or oracle/_ref (the byte-for-byte copy
`__graft_entry__.build()` stages with oracle/stage_ref.py so that it travels to the GPU box).  Used by
oracle/make_golden*.py to generate tests/golden/*.npz, by bench.py's `--impl reference` arm / `cpu_baseline` leg and
by tests/test_gpu_insitu.py.  Nothing on the product path (mvip_nerf_b200/) may import this module.

The reference's run.py imports six modules that do no arithmetic on the hot path
(matplotlib, imageio, tkinter, lpips, tinycudann, configargparse — SURVEY.md §0.4);
they are absent here, so empty stand-ins are registered before the import.
"""
import os
import sys
import types

def _find_root():
    cands = [os.environ.get("MVIP_REFERENCE_ROOT"), "/root/reference",
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "DS_NeRF", "run.py")):
            return c
    return cands[1]


REF_ROOT = _find_root()
REF_DIR = os.path.join(REF_ROOT, "DS_NeRF")

_STUBS = ["matplotlib", "matplotlib.pyplot", "imageio", "tkinter", "lpips",
          "tinycudann", "configargparse", "cv2", "torchvision", "skimage",
          "skimage.metrics", "scipy.spatial"]


def available():
    return os.path.isfile(os.path.join(REF_DIR, "run.py"))


def _stub(name):
    if name in sys.modules:
        return
    try:
        __import__(name)
        return
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package so "import a.b" works
    sys.modules[name] = m
    if "." in name:
        parent, child = name.rsplit(".", 1)
        _stub(parent)
        setattr(sys.modules[parent], child, m)


def load():
    """Returns (run, run_nerf_helpers) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_DIR)
    for n in _STUBS:
        _stub(n)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    try:
        helpers = importlib.import_module("run_nerf_helpers")
        run = importlib.import_module("run")
    except ImportError as e:  # one more missing non-arithmetic dependency: stub and retry
        missing = getattr(e, "name", None)
        if not missing:
            raise
        _stub(missing)
        return load()
    return run, helpers
