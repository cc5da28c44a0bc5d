"""Only the HBM-stage table of bench.py (sampling / compositing / ray / normal kernels at a 262,144-ray batch)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mvip_nerf_b200 import ops, _lib
if os.environ.get("MVIP_LIB"):          # experiment builds (scripts/build_variant.sh)
    _lib.LIB_PATH = os.path.abspath(os.environ["MVIP_LIB"])
out = bench.hbm_stage_times(ops, torch.device("cuda", 0), bench.peaks()["hbm_gbs"])
for k, v in out.items():
    print("%-26s %8.4f ms  %6.0f GB/s  %5.1f %%" % (k, v["ms_median"], v["gbs"], 100 * v["frac_of_hbm_peak"]))
print(json.dumps(out))
