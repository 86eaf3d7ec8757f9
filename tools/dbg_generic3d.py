import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tps_b200, oracle_api
from common import rel_l2, tgv_state, node_coords_from_mesh
PI=np.pi
eq=int(sys.argv[1]); vm=float(sys.argv[2])
m = tps_b200.cartesian_hex_mesh(4, 5, 3, lo=(-PI,)*3, hi=(PI,)*3)
phys = tps_b200.Physics.dry_air(eq, vm, 0.1)
U = tgv_state(node_coords_from_mesh(m["elem_xyz"], 3))
ys={}
for path in ("fast","generic"):
    os.environ["TPSB_PATH"]=path
    op = tps_b200.RhsOperator(m, order=3, physics=phys)
    ys[path]=op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    g=op.fields()[1].cpu().numpy(); ys[path+"_g"]=g
N=len(U)//5
print("3D eq",eq,"vm",vm,[f"{rel_l2(ys['generic'][k*N:(k+1)*N], ys['fast'][k*N:(k+1)*N]):.2e}" for k in range(5)], "grad", rel_l2(ys['generic_g'], ys['fast_g']))
