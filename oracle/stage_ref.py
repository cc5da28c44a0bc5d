"""TEST / BENCH INFRASTRUCTURE ONLY — stages the *unmodified* reference render path under oracle/_ref/.

The reference (MVIP-NeRF, DS_NeRF/) is a Python tree; /root/reference exists only in the build container.  So that the
GPU box can (a) time the reference's own CPU implementation as the `--impl reference` arm of bench.py and (b) run the
in-situ drop-in test (tests/test_gpu_insitu.py: the reference's train()/render_path() code calling OUR kernels),
`__graft_entry__.build()` copies the Python modules that `DS_NeRF/run.py` imports, byte for byte, to oracle/_ref/DS_NeRF/.
oracle/_ref/ is git-ignored (never part of the history) but not gpurun-ignored, so it travels like a built .so.
Nothing under mvip_nerf_b200/ may import from it.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MVIP_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
# run.py's own import closure inside the tree (discovered with sys.modules after `import run`, oracle/ref_import.py)
SUBDIRS = ["", "utils", "colmapUtils"]


def stage():
    """Copies DS_NeRF/{*.py, utils/*.py, colmapUtils/*.py}; returns the manifest (relative path -> sha256) or None."""
    src = os.path.join(SRC, "DS_NeRF")
    if not os.path.isfile(os.path.join(src, "run.py")):
        return None
    manifest = {}
    for sub in SUBDIRS:
        d = os.path.join(src, sub)
        for name in sorted(os.listdir(d)):
            if not name.endswith(".py") or " " in name:
                continue
            rel = os.path.join("DS_NeRF", sub, name)
            out = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(d, name), out)
            manifest[rel] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    m = stage()
    print("staged %d files under %s" % (len(m), DST) if m else "reference tree not present at %s" % SRC)
