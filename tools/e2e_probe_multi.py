"""ge_flat_forceatlas on an N-GPU context, one iteration per call, with the library's own laps
(GE_VERBOSE): where the per-call time of the bench's e2e leg goes.  python tools/e2e_probe_multi.py [N]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.load_package()
from graph_embed_b200 import capi, graphs
import scipy.sparse as sp
import torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
A = graphs.rgg(500_000, 10.0, seed=7)
n = A.shape[0]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
Ap = sp.csr_matrix((pin(A.data), pin(A.indices), pin(A.indptr)), shape=A.shape)
x = pin(capi.reference_uniform(23, n * 2).reshape(n, 2).copy())
ctx = capi.Context(devices=list(range(N)))
p1 = capi.flat_params(iterations=1)
ctx.flat_forceatlas(Ap, 2, x, p1)
for rep in range(8):
    if rep == 5:
        os.environ["GE_VERBOSE"] = "1"
    t = time.time()
    ctx.flat_forceatlas(Ap, 2, x, p1, inplace=True)
    print("call %d: %.1f ms" % (rep, 1e3 * (time.time() - t)), flush=True)
