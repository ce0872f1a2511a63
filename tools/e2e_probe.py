import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from amps_b200 import api, mesh as meshmod, workload
n=64; ppc=64
m = meshmod.uniform_periodic_box((n,n,n),(8,8,8),(1,1,1))
charge, mass, wgt = workload.species_tables(ppc, 1.0)
x,v,w,sp,cells = workload.maxwellian_box(m, ppc, seed=100)
cfg = api.make_config((8,8,8),(1,1,1),charge,mass,wgt,1.0,periodic=True,capacity=x.shape[1]+1024)
E,B = workload.box_fields(m); 
g = api.Context(cfg, m); g.fields_upload(E,B,B); g.particles_upload(x,v,w,sp,cells); g.sort()
Jh = torch.empty((m.n_corners,3),dtype=torch.float64).pin_memory().numpy()
Mh = torch.empty((m.n_corners,243),dtype=torch.float64).pin_memory().numpy()
Mp = np.empty((m.n_corners,243))
def t(f, k=3):
    f(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(k): f()
    g.synchronize(); return (time.perf_counter()-t0)/k*1e3
print("step", t(lambda: g.step()))
print("JM_download pinned", t(lambda: g.JM_download(out_J=Jh,out_M=Mh)))
print("step+download", t(lambda: (g.step(), g.JM_download(out_J=Jh,out_M=Mh))))
print("step_JM pinned", t(lambda: g.step_JM(Jh,Mh)))
print("step_JM pageable", t(lambda: g.step_JM(np.empty((m.n_corners,3)),Mp)))
