"""CPU suite: the C-ABI shared library loads and exports every symbol include/mvip_nerf.h declares
(no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mvip_nerf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvip_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from mvip_nerf_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names            # the ctypes table mirrors the header
    assert _lib.load().mvip_abi_version() == 1


def test_sizes_and_errors_without_gpu():
    from mvip_nerf_b200 import _lib
    lib = _lib.load()
    assert lib.mvip_mlp_packed_bytes() == 34 * 32768 + 5 * 16384 + 34 * 32768 + 3080 * 4
    assert lib.mvip_mlp_stash_bytes(129) == 2 * (40 * 16384 + 9 * 128 * 32)
    assert lib.mvip_normal_workspace_bytes(4, 5) == 2 * 9 * 20 * 8
    # argument validation happens before any CUDA call
    rc = lib.mvip_sample_coarse(None, 11, 4, None, None, 64, 1, None, None)   # n_rays=4 with null buffers
    assert rc == -1 and b"null" in lib.mvip_last_error()


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "mvip_nerf_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"(^|\n)\s*(from|import)\s+[\w., ]*oracle|#include[^\n]*oracle|oracle/_ref|dlopen", txt), f


def test_ops_refuse_cpu_tensors():
    import torch
    from mvip_nerf_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sample_coarse(torch.zeros(4, 11), torch.linspace(0, 1, 64))
