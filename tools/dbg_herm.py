import sys, time
sys.path[:0]=["/root/repo","/root/repo/direct-nbody_b200"]
import numpy as np, nbk, nbh
from oracle import oracle as O
m,q,p = O.rescale_to_standard_units(*O.make_plummer(64, 20240))
ctx = nbk.Context(0)
for eta in (1e-4, 1e-6):
    sa = nbh.HermiteState(m,q,p); sb = O.HermiteState(m,q,p)
    g, dg = ctx.grad_v_dgrad_v_sum(m,q,p,0.0)
    f, df = -g, -dg
    h = 6.0*eta*np.linalg.norm(f,axis=1)/(m*np.linalg.norm(df,axis=1))
    ctx.hermite_upload(np.zeros(64), m, q, p, f, df, h)
    for k in range(4):
        stop = 0.25*(k+1)
        t0=time.time()
        try:
            ns = ctx.hermite_advance_resident(stop, eta, 0.0, max_steps=400000)
        except Exception as e:
            print("ERR", e); ns=-1
        t,qq,pp,hh,ff,dff = ctx.hermite_download(64)
        sb = O.hermite_advance(sb, 0.25, eta)
        print(f"eta {eta} call {k}: nsteps {ns} ({time.time()-t0:.2f}s) oracle total {sb.nsteps}; t range {t.min()} {t.max()} h range {hh.min():.3e} {hh.max():.3e} nan {np.isnan(qq).any()} oracle h {sb.h.min():.3e} {sb.h.max():.3e}")
        if ns < 0: break
