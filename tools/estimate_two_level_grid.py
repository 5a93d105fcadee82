"""Estimate (CPU, numpy) of the AABB tests per volume a two-level grid would need on the clustered scene
(configs[3] generator at 400k colliders), against the single-level grid the kernels use.  DESIGN.md section 9.
Run from the repo root: python tools/estimate_two_level_grid.py   (uses the oracle only to build the AABBs)"""
import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from gunship_rs_b200 import scenes
import oracle_lib as O
n=400000
s=scenes.by_config(4,n)
sysm=O.OracleSystem(8,8)
sysm.bvh_update(s.entity,s.kind,s.offset,s.size,s.pos,s.quat,s.scale)
v=sysm.volumes()
mn=v['aabb_min'][:,:3].astype(np.float64); mx=v['aabb_max'][:,:3].astype(np.float64)
ext=(mx-mn).max(axis=1); cell=ext.max()
print('cell',cell,'ext percentiles',np.percentile(ext,[10,50,90,99]))
def count_tests(mn,mx,cell,mask_a,mask_b,half=True):
    # number of (a,b) candidate tests where a in A, b in B, home cells differ by <=1 per axis (half-shell if same set)
    ca=np.floor(mn[mask_a]/cell).astype(np.int64); cb=np.floor(mn[mask_b]/cell).astype(np.int64)
    off=min(ca.min(),cb.min())-1
    ca-=off; cb-=off
    dims=max(ca.max(),cb.max())+3
    kb=(cb[:,0]*dims+cb[:,1])*dims+cb[:,2]
    cnt=np.bincount(kb,minlength=dims**3)
    tot=0
    for dx in (-1,0,1):
        for dy in (-1,0,1):
            for dz in (-1,0,1):
                k=((ca[:,0]+dx)*dims+(ca[:,1]+dy))*dims+(ca[:,2]+dz)
                tot+=cnt[k].sum()
    if half: tot=(tot-mask_a.sum())/2   # unordered pairs within the same set
    return tot
allm=np.ones(n,bool)
base=count_tests(mn,mx,cell,allm,allm)
print('single level: tests per volume',base/n)
for frac in (0.5,0.33,0.25):
    small=ext<=cell*frac
    ns=small.sum()
    t_ss=count_tests(mn,mx,cell*frac,small,small)          # fine grid, both small: spans <=2 fine cells
    t_ll=count_tests(mn,mx,cell,~small,~small)
    t_sl=count_tests(mn,mx,cell,small,~small,half=False)     # each small looks at large in 27 coarse cells
    print('fine = cell*%.2f: small %.2f; tests/volume ss %.2f ll %.2f sl %.2f total %.2f'%(frac,ns/n,t_ss/n,t_ll/n,t_sl/n,(t_ss+t_ll+t_sl)/n))

def cross_tests(mn,mx,cell,small,large):
    # small A looks at large B with home cell in [cell(A.min)-1, cell(A.max)] per axis
    cb=np.floor(mn[large]/cell).astype(np.int64)
    a0=np.floor(mn[small]/cell).astype(np.int64)-1; a1=np.floor(mx[small]/cell).astype(np.int64)
    off=min(cb.min(),a0.min())-1
    cb-=off; a0-=off; a1-=off
    dims=max(cb.max(),a1.max())+3
    cnt=np.bincount((cb[:,0]*dims+cb[:,1])*dims+cb[:,2],minlength=dims**3).reshape(dims,dims,dims)
    # 3-D prefix sums for box queries
    P=np.zeros((dims+1,dims+1,dims+1),np.int64); P[1:,1:,1:]=cnt.cumsum(0).cumsum(1).cumsum(2)
    x0,y0,z0=a0[:,0],a0[:,1],a0[:,2]; x1,y1,z1=a1[:,0]+1,a1[:,1]+1,a1[:,2]+1
    q=(P[x1,y1,z1]-P[x0,y1,z1]-P[x1,y0,z1]-P[x1,y1,z0]+P[x0,y0,z1]+P[x0,y1,z0]+P[x1,y0,z0]-P[x0,y0,z0])
    return q.sum()
def same_tests_pruned(mn,mx,cell,m):
    # every ordered (A,B) with B.home in [cell(A.min)-1, cell(A.max)] per axis, halved (what reach pruning approaches)
    return (cross_tests(mn,mx,cell,m,m)-m.sum())/2
print('single level with reach-style pruning: %.2f'%(same_tests_pruned(mn,mx,cell,allm)/n))
for frac in (0.5,0.4,0.33):
    small=ext<=cell*frac
    t=same_tests_pruned(mn,mx,cell*frac,small)+same_tests_pruned(mn,mx,cell,~small)+cross_tests(mn,mx,cell,small,~small)
    print('fine = cell*%.2f: small %.2f: tests/volume %.2f'%(frac,small.mean(),t/n))
