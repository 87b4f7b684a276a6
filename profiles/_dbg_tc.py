import sys
sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/multiview-clustering_b200')
import numpy as np, pyoracle as po, mvc_b200
from conftest import make_mixture
n,k_true,V,cap=int(sys.argv[1]),int(sys.argv[2]),3,64
views,z=make_mixture(n,[64]*V,k_true,seed=11)
rng=np.random.default_rng(2)
tab=np.where(rng.random(n)<0.15,rng.integers(0,k_true,n),z).astype(np.int32)
tab[:3]=[k_true+1,k_true+2,k_true+3]
dish=np.full((V,cap),-1,np.int32)
for t in range(k_true+4):
    dish[:,t]=rng.integers(0,max(2,k_true-1),3)[:V]
s=mvc_b200.Sampler(n,[64]*V,cap=cap,seed=123,engine=2,debug_export=True)
for v in range(V): s.upload_view(v,views[v])
s.set_state(tab,dish,np.full(V,1.0),np.full(V,0.5),np.full(V,0.9),1.0,0.6,sweep=3)
pre=s.get_state(); P=s.get_params()
o=po.OracleState(views,cap,seed=123)
o.alpha_v[:]=pre["alpha_v"];o.sigma_v[:]=pre["sigma_v"];o.tau_v[:]=pre["tau_v"];o.alpha_g=pre["alpha_g"];o.sigma_g=pre["sigma_g"];o.sweep=pre["sweep"]
o.set_assignment(pre["table_of"],pre["dish_of"])
for v in range(V): o.S1[v][:]=pre["S1"][v]
o.S2[:]=pre["sum_y2"]
s.sweep(1,True)
acc,xx,raw=s.get_debug_rows(); lnew=s.get_debug_lnew()
ps=po.params_struct(P)
bad=[]
for i in range(0,n,7):
    lw64=o.row_logweights(i)/np.log(2.0)
    ch,lw32=po.stageB_tc(ps,acc[i],xx[i],pre["table_of"][i],0.5,lnew[i],want_lw=True)
    ok=np.isfinite(lw64)
    err=np.abs(lw32-lw64)/np.maximum(1.0,np.abs(lw64))
    err[~ok]=0
    j=int(np.argmax(err))
    if err[j]>1e-4: bad.append((i,j,int(pre["table_of"][i]),float(lw32[j]),float(lw64[j]),i%128,(i//128)))
print(len(bad),'bad of',len(range(0,n,7)))
for b in bad[:25]: print(b)
