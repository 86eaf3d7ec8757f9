"""Development helper: per-equation parity errors of the fast path vs the oracle and vs the legacy kernels."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tps_b200, oracle_api
from common import rel_l2, tgv_state
PI = np.pi
n3 = tuple(int(x) for x in (sys.argv[1:4] or (6, 6, 6)))
order = int(sys.argv[4]) if len(sys.argv) > 4 else 3
vm = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
eq = int(sys.argv[6]) if len(sys.argv) > 6 else 1
m = tps_b200.cartesian_hex_mesh(*n3, lo=(-PI,) * 3, hi=(PI,) * 3)
orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                        phys=oracle_api.dry_air_params(eq, vm, 0.3))
U = tgv_state(orc.node_coords())
yo = orc.mult(U)
N = orc.N
res = {}
for path in ("fast", "legacy"):
    os.environ["TPSB_PATH"] = path
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(eq, vm, 0.3))
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    res[path] = y
    res[path + "_fr"] = op.debug_buffer(0).cpu().numpy().copy()
    print(path, [f"{rel_l2(y[k*N:(k+1)*N], yo[k*N:(k+1)*N]):.2e}" for k in range(5)], "mcs", op.max_char_speed() / orc.max_char_speed - 1)
    op.close()
d = np.abs(res["fast"] - res["legacy"]).reshape(5, -1, (order + 1) ** 3)
print("fast vs legacy max abs per eq", d.max(axis=(1, 2)), "scale", np.abs(res["legacy"]).reshape(5, -1).max(axis=1))
bad = np.argwhere(d[4] > 1e-6 * np.abs(res["legacy"]).max())
print("bad elements (eq 4):", np.unique(bad[:, 0])[:20], "count", len(np.unique(bad[:, 0])), "of", d.shape[1])

fr_f, fr_l = res["fast_fr"].reshape(-1, 5, (order + 1) ** 2), res["legacy_fr"].reshape(-1, 5, (order + 1) ** 2)
df = np.abs(fr_f - fr_l).max(axis=(1, 2))
print("faceRes: faces differing", int((df > 1e-9 * np.abs(fr_l).max()).sum()), "of", len(df), " worst", df.max(), "scale", np.abs(fr_l).max())
w = int(np.argmax(df))
np.set_printoptions(linewidth=200, precision=6)
print("worst face", w, "el1/el2", m["face_el1"][w], m["face_el2"][w], "inf", m["face_inf1"][w], m["face_inf2"][w])
print("fast  eq0", fr_f[w, 0]); print("legacy eq0", fr_l[w, 0])
print("fast  eq4", fr_f[w, 4]); print("legacy eq4", fr_l[w, 4])
