"""Quick GPU check of the tensor-core TriDistance path against the CPU oracle (dev tool)."""
import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import geometrics_b200 as g
from helpers import tri_inputs, TRI_CASES
from oracle import Oracle
orc = Oracle(); dev = torch.device('cuda:0')
def cmp(name, pts, t1, t2, t3, **kw):
    got = g.tri_forward(*(t.to(dev) for t in (pts,t1,t2,t3)), **kw)
    torch.cuda.synchronize()
    wd,wp,wi = orc.tri_fwd(pts.numpy(),t1.numpy(),t2.numpy(),t3.numpy())
    gd,gp,gi = (t.cpu().numpy() for t in got)
    bad = (gi!=wi) | (gp!=wp) | ((gd.view(np.int32)!=wd.view(np.int32)) & ~(np.isnan(gd)&np.isnan(wd)))
    print(name, kw, "mismatches", int(bad.sum()), flush=True)
    for b_, i_ in zip(*np.nonzero(bad)):
        print("   at", b_, i_, "p", pts[b_, i_].tolist(), "got", gd[b_,i_], gp[b_,i_], gi[b_,i_], "want", wd[b_,i_], wp[b_,i_], wi[b_,i_])
for case in TRI_CASES:
    if case[3] >= 4: cmp(str(case), *tri_inputs(*case), tensor=True)
pts, t1, t2, t3 = tri_inputs("sphere", 2, 600, 960, 34)
pts[0, 5, 1] = float("nan"); pts[1, 17, 0] = float("inf"); pts[1, 300] = float("-inf")
cmp("nonfinite", pts, t1, t2, t3, tensor=True)
cmp("nonfinite", pts, t1, t2, t3, tensor=False)
