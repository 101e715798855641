import sys; sys.path.insert(0, "/root/repo")
import torch
from noa_b200 import dcs, grids, STANDARD_ROCK, MUON_MASS
K = torch.from_numpy(grids.table_energies(10000)).cuda()
r = torch.zeros_like(K)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); best=1e9
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b))
    return best
print("soft_scattering 1e4 energies: %.4f ms" % t(lambda: dcs.soft_scattering(r, K, STANDARD_ROCK, MUON_MASS)))
