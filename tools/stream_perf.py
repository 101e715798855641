import sys
sys.path.insert(0, "/root/repo")
import torch
from noa_b200 import dcs, grids, STANDARD_ROCK, MUON_MASS
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for n in (1<<22, 1<<24):
    K,q=grids.set_b(n); Kd,qd=torch.from_numpy(K).cuda(),torch.from_numpy(q).cuda(); r=torch.empty_like(Kd)
    for pr in dcs.PROCESSES:
        if n == 1<<24 and pr.index in (1,2): continue
        t=timeit(lambda: dcs.vmap(pr)(r,Kd,qd,STANDARD_ROCK,MUON_MASS))
        print(f"   n=2^{n.bit_length()-1} {pr.name:16s} {t:8.4f} ms {n/t/1e6:8.2f} G/s", flush=True)
# pinned (mapped) path
n=1<<22
K,q=grids.set_b(n); Kh,qh=torch.from_numpy(K).pin_memory(),torch.from_numpy(q).pin_memory(); oh=torch.empty(n,dtype=torch.float64).pin_memory()
st=dcs.HostStager()
import time
for pr in (dcs.bremsstrahlung, dcs.pair_production):
    st.map(pr,Kh,qh,STANDARD_ROCK,MUON_MASS,out=oh); ts=[]
    for _ in range(10):
        t0=time.perf_counter(); st.map(pr,Kh,qh,STANDARD_ROCK,MUON_MASS,out=oh); ts.append(time.perf_counter()-t0)
    print(f"   pinned {pr.name:16s} {min(ts)*1e3:8.4f} ms", flush=True)
