"""Build-container check (needs oracle/_ref/ref_harness, i.e. /root/reference): runs the unmodified reference on its run.sh case
for several (levelMax, steps) pairs and, on each resulting multi-level mesh, compares with the reference (1) the numpy AMR
oracle (operators, flux-corrected: bit-exact), (2) the ghost-stencil tables of the C++ plan applied to the fields (labs:
1e-13), (3) the native Poisson rows (bitwise the reference's COO).  Round-1 run: (8,12) 281 blocks, (8,30) 305, (9,2) 593,
(9,15) 587 blocks over up to 8 levels: all bit-exact / 1.3e-15 / bitwise, 76-100 table patterns, no fallbacks."""
import sys, subprocess, os, numpy as np, scipy.sparse as sp, collections
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'oracle'))
import cup2d_amr_oracle as A
from cup2d_b200.amr import AmrPlan, LAB_SHAPES
H=os.path.join(ROOT,'oracle','_ref','ref_harness')
def run(lmax, nsteps):
    out=f'/tmp/amr_{lmax}_{nsteps}.bin'; coo=f'/tmp/coo_{lmax}_{nsteps}.bin'
    subprocess.run([H,'amrlab',str(lmax),str(nsteps),out],check=True,stderr=subprocess.DEVNULL,stdout=subprocess.DEVNULL,env=dict(os.environ,OMP_NUM_THREADS='1',CUP2D_REF_DUMP_COO=coo))
    a=np.fromfile(out); i=0; rec={}
    while i<len(a):
        tag,n=int(a[i]),int(a[i+1]); rec[tag]=a[i+2:i+2+n]; i+=2+n
    nu,dt,h0,bpdx,bpdy,_=rec[10]; blocks=rec[11].reshape(-1,3).astype(np.int32); nb=len(blocks)
    mesh=A.Mesh(blocks,int(bpdx),int(bpdy))
    vel=rec[12].reshape(nb,8,8,2); pres=rec[13].reshape(nb,8,8,1); chi=rec[14].reshape(nb,8,8,1); udef=rec[15].reshape(nb,8,8,2)
    o=A.amr_operators(mesh,h0,vel,pres,chi,udef,nu,dt)
    ok={k:bool(np.array_equal(o[k],rec[t].reshape(o[k].shape))) for k,t in (("adv",30),("rhs",31),("rhs1",32),("gradp",33))}
    plan=AmrPlan(blocks,int(bpdx),int(bpdy))
    worst=0
    for which,tag,f in ((0,20,vel),(1,21,vel),(2,22,pres)):
        rp,sb,sc,w=plan.stencil(which); ny,nx,dim=LAB_SHAPES[which]
        T=sp.csr_matrix((w,sb.astype(np.int64)*(64*dim)+sc,rp),shape=(nb*ny*nx*dim,nb*64*dim))
        lab=(T@f.reshape(-1)); g=rec[tag]; written=np.diff(rp)>0
        worst=max(worst,np.abs(lab[written]-g[written]).max()/np.abs(g).max())
    plan.ghosts(0); st=plan.stats(0)
    # poisson
    raw=open(coo,'rb').read(); m,nnz=np.frombuffer(raw,dtype=np.int64,count=2)
    same_mesh = m==64*nb
    pois=None
    if same_mesh:
        r=np.frombuffer(raw,dtype=np.int32,count=nnz,offset=16); c=np.frombuffer(raw,dtype=np.int32,count=nnz,offset=16+4*nnz); v=np.frombuffer(raw,dtype=np.float64,count=nnz,offset=16+8*nnz)
        ref=sp.coo_matrix((v,(r,c)),shape=(m,m)).tocsr(); ref.sort_indices()
        nbr,rows,rowptr,col,val=plan.poisson()
        pois=True
        for q,rr in enumerate(rows):
            a0,b0=ref.indptr[rr],ref.indptr[rr+1]; nz=ref.data[a0:b0]!=0; mz=val[rowptr[q]:rowptr[q+1]]!=0
            if not (np.array_equal(ref.indices[a0:b0][nz],col[rowptr[q]:rowptr[q+1]][mz]) and np.array_equal(ref.data[a0:b0][nz],val[rowptr[q]:rowptr[q+1]][mz])): pois=False; break
    print(lmax,nsteps,"blocks",nb,"levels",sorted(collections.Counter(blocks[:,0]).items()),"ops bit-exact",ok,"tables rel err %.1e"%worst,st,"poisson rows",pois)
    plan.close()
for lm,ns in ((8,12),(8,30),(9,2),(9,15)):
    try: run(lm,ns)
    except Exception as e: print(lm,ns,"FAILED",repr(e)[:300])
