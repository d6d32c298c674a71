import sys, heapq, numpy as np
sys.path.insert(0, '.')
from oracle.skimage_shim.segmentation import watershed
def minimax_forest(image, markers, mask, iters=None):
    """Jacobi fixed point of: q* = argmin over masked 4-neighbours of (b, d, label); b = max(v, b_q*), d = d_q*+1, label = label_q*"""
    h,w=image.shape
    INF=np.inf
    src = (markers>0) & mask
    b=np.where(src, image.astype(np.float64), INF); d=np.where(src,0,1<<30).astype(np.int64); lab=np.where(src,markers,0).astype(np.int64)
    free = mask & ~src
    n=0
    while True:
        n+=1
        best=None
        for dy,dx in ((-1,0),(0,-1),(0,1),(1,0)):
            bq=np.full((h,w),INF); dq=np.full((h,w),1<<30,dtype=np.int64); lq=np.zeros((h,w),dtype=np.int64)
            ys=slice(max(dy,0),h+min(dy,0)); xs=slice(max(dx,0),w+min(dx,0))
            yd=slice(max(-dy,0),h+min(-dy,0)); xd=slice(max(-dx,0),w+min(-dx,0))
            bq[yd,xd]=b[ys,xs]; dq[yd,xd]=d[ys,xs]; lq[yd,xd]=lab[ys,xs]
            if best is None: best=(bq,dq,lq)
            else:
                B,D,L=best
                better=(bq<B)|((bq==B)&((dq<D)|((dq==D)&(lq<L))))
                best=(np.where(better,bq,B),np.where(better,dq,D),np.where(better,lq,L))
        B,D,L=best
        reach=np.isfinite(B)&free
        nb=np.where(reach,np.maximum(image,B),b); nd=np.where(reach,D+1,d); nl=np.where(reach,L,lab)
        nb=np.where(free&~reach,INF,nb); nl=np.where(free&~reach,0,nl)
        if np.array_equal(nb,b) and np.array_equal(nd,d) and np.array_equal(nl,lab): break
        b,d,lab=nb,nd,nl
    return lab.astype(np.int32), n
rng=np.random.default_rng(0)
bad=0
for t in range(200):
    h,w=rng.integers(4,40,2)
    img=rng.permutation(h*w).reshape(h,w).astype(np.float32)   # tie-free
    if t%3==0:  # smooth-ish structure
        yy,xx=np.mgrid[0:h,0:w]; img=(np.sin(yy/3.0)+np.cos(xx/4.0))*50+img/(h*w)
        img=img.astype(np.float64)
    mask=rng.random((h,w))<rng.uniform(0.6,1.0)
    markers=np.zeros((h,w),np.int32)
    k=rng.integers(1,8)
    for i in range(k):
        y,x=rng.integers(0,h),rng.integers(0,w); markers[y:y+rng.integers(1,3),x:x+rng.integers(1,3)]=i+1
    want=watershed(img,markers,mask=mask)
    got,n=minimax_forest(img,markers,mask)
    if not np.array_equal(got,want):
        bad+=1; print("MISMATCH",t,h,w,(got!=want).sum())
print("done bad=",bad)
