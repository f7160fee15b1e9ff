import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g
pkg = g.load_pkg(); W = pkg.workloads
ctx = pkg.default_context()
os.environ["GMRFB_BTD_LOOKAHEAD"] = "0"
D, Bs = W.random_btd(512, 3, seed=1)
F = pkg.tridiagonal_cholesky_dense(D, Bs, ctx=ctx)
print(F.logdet())
