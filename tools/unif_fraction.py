"""Offline estimate (CPU, numpy) of how many a-trous tiles / warps of the 4K benchmark scene see one normal vector, per dilation:
the numbers quoted in DESIGN.md section 6 and 10 for the uniform-normal shortcut.  python tools/unif_fraction.py"""
import numpy as np, sys, time
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from svgf_b200 import synth
W,H=3840,2160
t0=time.time()
pl=synth.frame_host(W,H,10)
print("gen",time.time()-t0, {k:(v.shape,v.dtype) for k,v in pl.items()})
n=pl["normal"].view(np.uint16)[...,:3].astype(np.uint64)
key=(n[...,0]|(n[...,1]<<16)|(n[...,2]<<32))
z=pl["motion"][...,2]
bg=(z==0)
print("bg frac",bg.mean())
def region_uniform(key,valid,y0,y1,ys,x0,x1):
    ys_=np.arange(y0,y1,ys); ys_=ys_[(ys_>=0)&(ys_<H)]
    xa,xb=max(x0,0),min(x1,W)
    if len(ys_)==0 or xa>=xb: return True,False
    k=key[ys_][:,xa:xb]; v=valid[ys_][:,xa:xb]
    kk=k[v]
    if kk.size==0: return True,False
    return bool((kk==kk[0]).all()), True
inimg=np.ones((H,W),bool)
for step in (1,2,4,8,16):
    tiles=0; uni_tile=0; warps=0; uni_warp_old=0; uni_warp=0; uni_warp_bg=0
    TR=12
    rng=np.random.default_rng(0)
    ntx=W//128; nty=((H+TR*step-1)//(TR*step))*step
    sample=[(rng.integers(ntx),rng.integers(nty)) for _ in range(400)]
    for bx,by in sample:
        x0=bx*128; yb=by//step; ph=by%step; y0=yb*TR*step+ph
        # tile rule (current): all in-image texels incl. bg must equal first pixel and first pixel nonzero
        u,_=region_uniform(key,inimg,y0-2*step,y0+(TR+2)*step,step,x0-2*step,x0+128+2*step)
        first=key[min(y0,H-1),x0]
        ut = u and first!=0
        tiles+=1; uni_tile+=ut
        for w in range(2):
            for tg in range(4):
                ry0=y0+(tg*3-2)*step; ry1=y0+(tg*3+5)*step
                # conservative block cols: staged pair cols [32w, 32w+64) -> pixels x0-2step+64w .. +128
                xa=x0-2*step+64*w; xb=xa+128
                xb=min(xb, x0+128+2*step)
                warps+=1
                if ut: uni_warp_old+=1
                u1,_=region_uniform(key,inimg,ry0,ry1,step,xa,xb)
                k0=key[min(max(y0+tg*3*step,0),H-1),x0+64*w]
                uni_warp += (u1 and k0!=0)
                u2,any_=region_uniform(key,~bg,ry0,ry1,step,xa,xb)
                uni_warp_bg += (u2 and any_)
    print(step,"tile unif %.2f | warp: old %.2f new %.2f new+bgexcl %.2f"%(uni_tile/tiles,uni_warp_old/warps,uni_warp/warps,uni_warp_bg/warps))
