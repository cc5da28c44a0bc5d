"""In-situ drop-in test (VERDICT r1 missing #2): the reference's own train()-iteration render calls, render_path and
render_path_4view run once with the stock reference functions (fp32 PyTorch on the B200) and once with the names rebound
exactly as INTEGRATION.md §1 prescribes, on the same inputs and the same CUDA random stream.  tests/insitu_worker.py does the
work in a child process (the reference flips torch's default tensor type, DS_NeRF/run.py:1978)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_reference_train_iteration_and_render_path_on_our_kernels():
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not staged (oracle/_ref is written by __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "insitu_worker.py")], capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-3000:])
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("INSITU ")][-1]
    r = json.loads(line[7:])
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(r, open(os.path.join(out, "insitu.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in r.items() if k != "grad"}, indent=1))
    assert r["launches"] > 50 and r["files_equal"] and r["n_files"] >= 2 * 7       # same on-disk layout from the reference's writer
    # integrated maps, ALL rays (random-init net, perturb + noise on: the far-sample conditioning caveat applies to a few rays,
    # hence mean-abs bounds beside generous max-abs ones); tolerances = DESIGN.md §2, measured values in profiles/r02_parity.md
    for k in ("rgb", "rgb2", "rgb0", "rgbs4", "path_rgbs"):
        assert r[k]["mean_abs"] <= 1e-3 and r[k]["nan_mismatch"] == 0, (k, r[k])
    for k in ("disp", "disp2", "path_disps"):
        assert r[k]["mean_abs"] <= 5e-3 * max(1.0, r[k]["ref_mean"]), (k, r[k])
    assert r["depth1"]["mean_abs"] <= 1e-2 * r["depth1"]["ref_mean"], r["depth1"]
    assert abs(r["loss_ours"] - r["loss_ref"]) <= 2e-3 * abs(r["loss_ref"]), (r["loss_ours"], r["loss_ref"])
    assert abs(r["loss_after_ours"] - r["loss_after_ref"]) <= 5e-3 * abs(r["loss_after_ref"])
    assert r["loss_after_ours"] < r["loss_ours"] * 1.5
    assert r["grad_min_cos"] >= 0.98, r["grad_min_cos"]
