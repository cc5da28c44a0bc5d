"""mvip_nerf_b200 — B200-native (sm_100a) implementation of MVIP-NeRF's NeRF volume-rendering hot path.

Drop-in names (same signatures as the reference's DS_NeRF/run.py and run_nerf_helpers.py) live in
`mvip_nerf_b200.run_nerf_helpers` and `mvip_nerf_b200.run`; the kernels are reached through the C ABI of
libmvip_nerf.so (include/mvip_nerf.h) via `mvip_nerf_b200.ops`.
"""
from . import _lib, ops  # noqa: F401

__all__ = ["ops", "_lib"]
