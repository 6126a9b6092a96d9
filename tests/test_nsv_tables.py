"""Host-side emulation of the range-minimum tables and the `first body >= j0 whose shared-level byte is
<= l` query that cells_kernel uses for the end of a cell's run (gravity.cu: unit_kernel, nsv_level2_kernel,
nsv_descend, nsv_next_le), checked against brute force for block / super-block edge sizes."""
import numpy as np
def build(A):
    n=len(A); nb=(n+255)//256; n_pad=nb*256
    t1=np.full((9,n_pad),255,np.uint8); t1[0,:n]=A
    for blk in range(nb):
        tab=np.full(384,255,np.uint8); tab[:256]=t1[0,blk*256:(blk+1)*256]
        for k in range(1,9):
            new=np.full(384,255,np.uint8)
            new[:256]=np.minimum(tab[:256],tab[(1<<(k-1)):(1<<(k-1))+256])
            t1[k,blk*256:(blk+1)*256]=new[:256]; tab=new
    nblocks=nb; b_pad=(nblocks+255)//256*256; nsuper=b_pad//256
    t2=np.full((9,b_pad),255,np.uint8); t2[0,:nblocks]=t1[8,::256][:nblocks]
    t3=np.full(nsuper,255,np.uint8)
    for sb in range(nsuper):
        tab=np.full(384,255,np.uint8); tab[:256]=t2[0,sb*256:(sb+1)*256]
        for k in range(1,9):
            new=np.full(384,255,np.uint8)
            new[:256]=np.minimum(tab[:256],tab[(1<<(k-1)):(1<<(k-1))+256])
            t2[k,sb*256:(sb+1)*256]=new[:256]; tab=new
        t3[sb]=tab[0]
    return dict(t1=t1,t2=t2,t3=t3,n=n,nblocks=nblocks,nsuper=nsuper)
def descend(tab,j,end,l):
    for k in range(8,-1,-1):
        step=1<<k
        if j+step<=end and tab[k,j]>l: j+=step
    return j
def next_le(tv,j0,l):
    n=tv['n']
    if j0>=n: return n
    blk_end=min(n,(j0|255)+1)
    j=descend(tv['t1'],j0,blk_end,l)
    if j<blk_end: return j
    if blk_end==n: return n
    b=blk_end>>8
    if b>=tv['nblocks']: return n
    sb_end=min(tv['nblocks'],(b|255)+1)
    b=descend(tv['t2'],b,sb_end,l)
    if b==sb_end:
        if sb_end==tv['nblocks']: return n
        q=sb_end>>8
        while q<tv['nsuper'] and tv['t3'][q]>l: q+=1
        if q>=tv['nsuper']: return n
        b=q<<8; sb_end=min(tv['nblocks'],b+256)
        b=descend(tv['t2'],b,sb_end,l)
        if b==sb_end: return n
    lo=b<<8
    return descend(tv['t1'],lo,min(n,lo+256),l)
def test_nsv_query_matches_brute_force():
  rng=np.random.default_rng(0)
  for n in [1,2,67,255,256,257,511,512,1000,65535,65536,65537,70000,131072+300, 200000]:
      # mostly large values with rare small ones so answers are far away
      A=rng.integers(5,40,n).astype(np.uint8)
      A[rng.random(n)<0.3]=255
      for pos in rng.integers(0,n,max(1,n//2000)): A[pos]=rng.integers(0,6)
      tv=build(A)
      for _ in range(400):
          j0=int(rng.integers(0,n+1)); l=int(rng.integers(0,42))
          idx=np.nonzero(A[j0:]<=l)[0]
          want=j0+int(idx[0]) if len(idx) else n
          got=next_le(tv,j0,l)
          assert got==want,(n,j0,l,got,want)
