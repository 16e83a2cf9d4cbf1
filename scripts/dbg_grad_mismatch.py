import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import __graft_entry__ as g
pkg=g.load_package()
from oracle import oracle as orc
from test_gpu_parity import setup, gpu_rollout
from conftest import oracle_problem
wl, sur, rn, starts, dd = setup(pkg, orc, "C2", M=128)
P = oracle_problem(orc, wl, sur, rn, starts, 1, dual_dirs=dd)
ref = P.rollout()
got = gpu_rollout(pkg, wl, sur, rn, starts, dd)
gscale = np.maximum(np.abs(ref["grad_x"]).max(axis=0, keepdims=True), 1e-6)
gerr = np.max(np.abs(got["grad_x"] - ref["grad_x"]) / gscale, axis=0)
bad = np.nonzero(gerr > 1e-6)[0]
print("bad", bad, gerr[bad])
for m in bad:
    print("m", m, "case", ref['grad_case'][m], got['grad_case'][m], "t", ref['best_index'][m], "xerr", np.abs(got['xs'][:,:,m]-ref['xs'][:,:,m]).max())
    print("  ref g", ref['grad_x'][:,m]); print("  got g", got['grad_x'][:,m])
    for j in range(1, ref['best_index'][m]+1):
        e = P.eval_point(ref['xs'][:,j,m], ref['xs'][:,:j,m], ref['ys'][:j,m])
        print("   step", j, "det Href", np.linalg.det(e['Halpha_ref']), "sigma", e['sigma'], "alpha", e['alpha'])
