"""Event timeline of one iteration of the forward pair kernel (library built with EXTRA=-DMVIP_TRACE)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvip_nerf_b200 import ops, _lib
from oracle import nerf_oracle as orc
if os.environ.get("MVIP_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MVIP_LIB"])
dev = "cuda"
p = orc.init_params(1)
blob = ops.mlp_pack([torch.from_numpy(p[n]).to(dev) for n in ops.PARAM_ORDER])
P = 128 * 4 * 74 * 8
pts = torch.rand(P, 3, device=dev) * 4 - 2
dirs = torch.nn.functional.normalize(torch.randn(P, 3, device=dev), dim=-1)
stash = len(sys.argv) > 1 and sys.argv[1] == "train"
for _ in range(3): ops.mlp_forward(blob, pts=pts, dirs=dirs, want_stash=stash)
torch.cuda.synchronize()
out = (ctypes.c_longlong * (3 * 1024 * 2))()
n = (ctypes.c_int * 3)()
rc = _lib.load().mvip_debug_trace(out, n)
assert rc == 0, rc
a = np.frombuffer(out, dtype=np.int64).reshape(3, 1024, 2)
ev = []
for role in range(3):
    for i in range(n[role]):
        ev.append((int(a[role, i, 1]), role, int(a[role, i, 0])))
ev.sort()
t0 = ev[0][0]
ISS = {0: "act-wait>", 1: "act-wait<", 2: "hi-wait>", 3: "hi-wait<", 4: "issued", 5: "weights ok"}
EPI = {0: "h0 acc-wait>", 1: "h0 wake", 2: "h0 loaded", 3: "h0 arrived(D free)", 4: "h1 acc-wait>", 5: "h1 wake", 6: "h1 loaded+st lo", 7: "h1 arrived(A lo)", 9: "h1 st hi done"}
for t, role, code in ev:
    if role == 0:
        s, h, k = code // 64, (code // 32) & 1, code & 31
        print("%7d  ISSUER-X s=%d h=%d %s" % (t - t0, s, h, ISS.get(k, str(k))))
    else:
        s, k = code // 16, code & 15
        print("%7d  %s EPI%d    s=%d %s" % (t - t0, "          " * role, role - 1, s, EPI.get(k, str(k))))
