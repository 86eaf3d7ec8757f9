import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tps_b200, oracle_api
from common import rel_l2
from test_gpu_generic_parity import _state2d
PI=np.pi
order,bt,ir,eq = [int(a) for a in sys.argv[1:5]]
vm = float(sys.argv[5]) if len(sys.argv)>5 else 3e4
m = tps_b200.cartesian_quad_mesh(7, 6, lo=(-PI,-PI), hi=(PI,PI))
op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(eq, vm, 0.2), basis_type=bt, int_rule_type=ir)
orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"], phys=oracle_api.dry_air_params(eq, vm, 0.2), basis_type=bt, int_rule=ir)
U=_state2d(orc.node_coords()); N=orc.N
y=op.Mult(torch.from_numpy(U).cuda()).cpu().numpy(); yo=orc.mult(U)
print([f"{rel_l2(y[k*N:(k+1)*N], yo[k*N:(k+1)*N]):.2e}" for k in range(4)])
d=np.abs(y-yo).reshape(4,-1,(order+1)**2)
print("per-eq max abs", d.max(axis=(1,2)), "scale", np.abs(yo).reshape(4,-1).max(axis=1))
print("worst elem per eq", d.max(axis=2).argmax(axis=1))
print("elem 0 eq0 ours", y[:N].reshape(-1,(order+1)**2)[0]); print("elem 0 eq0 orc ", yo[:N].reshape(-1,(order+1)**2)[0])
