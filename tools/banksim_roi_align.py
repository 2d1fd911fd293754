import numpy as np, torch, sys
sys.path.insert(0,'.')
from vision_b200 import workloads
x, rois, kw = workloads.cfg2_roi_align()
rois=rois.numpy().astype(np.float32); H,W=200,272; PH=PW=7; SR=2
def axis(v,size):
    v=np.maximum(v,0); lo=np.minimum(v.astype(np.int32), size-1); 
    lo=np.where(lo>=size-1, size-2, lo)   # 'lo=size-2,l=1' trick
    return lo
K=len(rois)
sw=rois[:,1]*0.25; sh=rois[:,2]*0.25; ew=rois[:,3]*0.25; eh=rois[:,4]*0.25
rw=np.maximum(ew-sw,1); rh=np.maximum(eh-sh,1); bw=rw/PW; bh=rh/PH
# sample coords [K, 14]
ys=sh[:,None,None]+ (np.arange(PH)[None,:,None]*bh[:,None,None]) + ((np.arange(SR)[None,None,:]+.5)*bh[:,None,None]/SR)
xs=sw[:,None,None]+ (np.arange(PW)[None,:,None]*bw[:,None,None]) + ((np.arange(SR)[None,None,:]+.5)*bw[:,None,None]/SR)
ylo=axis(ys,H); xlo=axis(xs,W)   # [K,7,2]
def wavefronts(addr):   # addr [nwarps,32] int (-1 inactive); returns total wavefronts
    tot=0
    for a in addr:
        a=a[a>=0]; 
        if len(a)==0: continue
        u=np.unique(a); b=u%32
        tot+=np.bincount(b,minlength=32).max()
    return tot
def sim(pitch, nthreads, order='pw'):
    nb=49
    # items: flat (roi, bin) ; thread t handles items t, t+nthreads...
    items=K*nb
    rng=np.arange(items)
    tot=0; ideal=0
    # sample a subset of iterations for speed
    its=range(0, (items+nthreads-1)//nthreads, 3)
    for it in its:
        idx=it*nthreads+np.arange(nthreads); valid=idx<items
        idx=np.where(valid,idx,0)
        n=idx//nb; b=idx%nb
        if order=='pw': ph=b//PW; pw=b%PW
        else: pw=b//PH; ph=b%PH
        for iy in range(SR):
            for ix in range(SR):
                base=ylo[n,ph,iy]*pitch+xlo[n,pw,ix]
                for d in (0,1,pitch,pitch+1):
                    a=np.where(valid,base+d,-1)
                    pad=(-len(a))%32
                    a=np.concatenate([a,-np.ones(pad,dtype=a.dtype)]).reshape(-1,32)
                    tot+=wavefronts(a); ideal+=(a>=0).any(1).sum()
    return tot/ideal
for pitch in (272,276,273,277,280,288+1):
    print('pitch',pitch,'1024thr', round(sim(pitch,1024),2), '980thr', round(sim(pitch,980),2))
print('ph-fastest, pitch 276, 980:', round(sim(276,980,'ph'),2))
