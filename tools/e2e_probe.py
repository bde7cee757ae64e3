import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
entry.load_package()
from graph_embed_b200 import capi, graphs
import torch
A = graphs.rgg(500_000, 10.0, seed=7)
n = A.shape[0]
x0 = capi.reference_uniform(23, n * 2).reshape(n, 2)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
import scipy.sparse as sp
Ap = sp.csr_matrix((pin(A.data), pin(A.indices), pin(A.indptr)), shape=A.shape)
x = pin(x0.copy())
ctx = capi.Context(0)
p1 = capi.flat_params(iterations=1)
for rep in range(3):
    ctx.flat_forceatlas(Ap, 2, x, p1)
os.environ["GE_VERBOSE_PLAN"] = "1"
t = time.time(); ctx.flat_forceatlas(Ap, 2, x, p1, inplace=True); print("call %.1f ms" % (1e3 * (time.time() - t)))
del os.environ["GE_VERBOSE_PLAN"]
for rep in range(3):
    t = time.time(); ctx.flat_forceatlas(Ap, 2, x, p1, inplace=True); print("call %.1f ms" % (1e3 * (time.time() - t)))
