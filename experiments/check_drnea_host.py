"""Multi-chunk check of the host entry points of the inverse-dynamics derivatives (rbd_rnea_derivatives_host / _host_sparse):
100 003 Talos instances (four chunks of the pipeline, a partial last tile), dense against support-only values bit for bit,
samples on the chunk boundaries against the oracle.  python experiments/check_drnea_host.py   (needs a GPU)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sofa.rigidbodydynamics_b200 import api, model as mdl
import oracle, time
d = mdl.talos_reduced(); n = 100003; nv = d.nv
q, v, a = mdl.random_state(d, n, seed=5)
m = api.Model(d); ws = api.Workspace(m, n)
hq = np.zeros((n, nv, nv)); hv = np.zeros((n, nv, nv)); hM = np.zeros((n, nv*(nv+1)//2))
t0=time.time(); ws.rnea_derivatives_host(n, q, v, a, hq, hv, hM); t1=time.time()
rp, ci = m.drnea_pattern(); nnz=len(ci)
sq = np.zeros((n, nnz)); sv = np.zeros((n, nnz))
t2=time.time(); ws.rnea_derivatives_host_sparse(n, q, v, a, sq, sv); t3=time.time()
rows = np.repeat(np.arange(nv), np.diff(rp))
assert np.array_equal(sq, hq[:, rows, ci]) and np.array_equal(sv, hv[:, rows, ci])
orc = oracle.Oracle(d)
for i in (0, 32767, 32768, 65535, 65536, n-1):
    rq, rv, rM = orc.rnea_derivatives(q[i], v[i], a[i])
    e = max(np.abs(hq[i]-rq).max()/max(1,np.abs(rq).max()), np.abs(hv[i]-rv).max()/max(1,np.abs(rv).max()))
    assert e < 1e-10, (i, e)
print("host derivatives ok: dense %.3f s (%.2f M/s), sparse %.3f s (%.2f M/s)"%(t1-t0, n/(t1-t0)/1e6, t3-t2, n/(t3-t2)/1e6))
