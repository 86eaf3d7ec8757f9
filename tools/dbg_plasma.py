import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tps_b200, oracle_api, plasma_cases
from common import rel_l2
from test_gpu_plasma import _pair
op, orc = _pair(n=(6,5))
xy = orc.node_coords(); up = plasma_cases.smooth_primitives(xy)
U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1); N = orc.N
y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy(); yo = orc.mult(U)
for k in range(6):
    a, b = y[k*N:(k+1)*N], yo[k*N:(k+1)*N]
    print(k, "nan count", np.isnan(a).sum(), "rel", rel_l2(np.nan_to_num(a), b))
bad = np.nonzero(np.isnan(y[4*N:5*N]))[0]
print("bad nodes", bad[:10], "of", N)
if len(bad):
    n = bad[0]
    Upd, gd = op.fields(); Upd = Upd.cpu().numpy().reshape(6, N); gd = gd.cpu().numpy().reshape(2, 6, N)
    print("Up at bad", Upd[:, n]); print("U at bad", U.reshape(6, N)[:, n]); print("grad at bad", gd[:, :, n])
    Un = U.reshape(6,N)[:, n][None,:].copy(); g = gd[:, :, n].reshape(1, 12).copy()
    print("oracle src", orc.pt("source", Un, orc.pt("prim", Un), g))
    print("device src", op.point_eval("source", torch.from_numpy(Un).cuda(), torch.from_numpy(g).cuda()).cpu().numpy())
