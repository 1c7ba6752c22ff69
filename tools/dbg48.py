import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from oracle import Oracle
from conftest import rel_err
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
for N,batch in ((48,40),(64,48),(40,30)):
    o=Oracle(N); rng=np.random.default_rng(N)
    K,F,Mt,fb=o.generate_rods(77,0,batch)
    q0=rng.normal(size=(batch,4)); q0/=np.linalg.norm(q0,axis=1,keepdims=True)
    r0=rng.normal(size=(batch,3)); lbar=rng.normal(size=(batch,3,N))
    ref=o.integrate_all(K,F,Mt,q0=q0,r0=r0,fbar=fb,lbar=lbar)
    ref2=o.integrate_all(K,F,Mt,q0=q0,r0=r0,fbar=fb,lbar=lbar,explicit_inverse=False)
    print(N,'oracle inv vs lu', {s: rel_err(ref[s],ref2[s]) for s in 'Qrnm'})
    t=lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    with SpectralRodIntegrator(N,0) as h:
        out=h.integrate_all(t(K),t(F),t(Mt),q0=t(q0),r0=t(r0),fbar=t(fb),lbar=t(lbar)); h.synchronize()
        print(N,'fused vs inv', {s: rel_err(out[s].cpu().numpy(),ref[s]) for s in 'Qrnm'})
        print(N,'fused vs lu ', {s: rel_err(out[s].cpu().numpy(),ref2[s]) for s in 'Qrnm'})
        Q=h.integrate_quaternions(t(K),q0=t(q0)); r=h.integrate_position(Q,r0=t(r0)); n=h.integrate_stress(t(F),fbar=t(fb)); m=h.integrate_couple(Q,n,t(Mt),q0=t(q0),lbar=t(lbar)); h.synchronize()
        print(N,'staged vs inv', {k: rel_err(v.cpu().numpy(),ref[k]) for k,v in (('Q',Q),('r',r),('n',n),('m',m))})
