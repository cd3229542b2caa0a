#!/usr/bin/env python
"""CPU study (no GPU): is the calibration of the tensor path's product count (hafgpu.cu, calibrate_tensor_passes)
representative?  It measures the operand error of the one- and two-product schemes with the model's own support vectors
standing in for the windows; here the same errors are measured on real windows (bundled clouds through the oracle's
bit-exact feature / scaling stages) for the trained substitute and for synthetic models of several gammas.
Result (profiles/r1_s2_calibration_study.txt): for the trained model the probes over-estimate (its SVs come from the
windows' own distribution); for synthetic SVs they under-estimate by at most 1.5x -- inside the 15x margin of guard_rel.

  python tools/calibration_study.py        (a few minutes on the CPU)
"""
import sys, os, gzip, tempfile, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haf_grasping_b200 import synth
from tools.dec_error_probe import load_model
from oracle import orc
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F=os.path.join(ROOT,"tests","golden","refdata","Features.txt"); R=os.path.join(ROOT,"tests","golden","refdata","range21062012_allfeatures")
def q16(a): return torch.from_numpy(np.ascontiguousarray(a,np.float32)).to(torch.float16).to(torch.float64).numpy()
tmp=tempfile.mkdtemp()
tm=os.path.join(tmp,'t.model'); open(tm,'wb').write(gzip.open(ROOT+'/tests/golden/substitute_trained.model.gz','rb').read())
orc.build(ref=False)
clouds=np.load(ROOT+'/tests/golden/clouds.npz')
o=orc.Oracle(F,R,tm)
rows=[]
for cn in ('table1','pcd2','table3'):
    ores=o.search(clouds[cn], orc.make_request())
    for roll in (0,5):
        feats,_=o.calc_featurevectors(ores['integral'][roll], ores['mask'][roll]); rows.append(o.scale(feats))
scaled=np.concatenate(rows); print('windows',len(scaled))
def rels(model, X):
    gamma,rho,coef,sv=model
    D=max(sv.shape[1],X.shape[1]); x=np.zeros((len(X),D)); x[:,:X.shape[1]]=X; s=np.zeros((len(sv),D)); s[:,:sv.shape[1]]=sv
    xh=q16(x); sh=q16(s); base=(x*x).sum(1)[:,None]+(s*s).sum(1)[None,:]
    dec=lambda dot: np.exp(-gamma*np.maximum(base-2*dot,0))@coef
    K=np.exp(-gamma*np.maximum(base-2*x@s.T,0)); E=(K*(1+gamma*1.4426950408889634*base))@np.abs(coef)+abs(rho)
    d=dec(x@s.T)
    return (np.abs(dec(xh@sh.T)-d)/E).max(), (np.abs(dec(xh@s.T)-d)/E).max()
for name,mp in [('trained',tm)]+[('synth2048 g=%g'%g, synth.write_synth_model(os.path.join(tmp,'s%g.model'%g),2048,gamma=g)) for g in (1/323.,0.005,0.01,0.02)]+[('synth512 g=1/323', synth.write_synth_model(os.path.join(tmp,'s512.model'),512))]:
    m=load_model(mp); S=len(m[3]); npb=min(S,48); idx=[k*S//npb for k in range(npb)]
    p1,p2=rels(m, m[3][idx]); r1,r2=rels(m, scaled)
    print('%-20s probes: rel1 %.2e rel2 %.2e | real windows: rel1 %.2e rel2 %.2e'%(name,p1,p2,r1,r2))
