"""Randomised edge-case sweep of the native kernels on the CPU SIMT emulator (tools/simt_emu): ragged lengths, channel counts
that are not multiples of the tile, groups, every dtype / direction / optional operand, compared against the oracle.

usage: python tools/emu_fuzz.py scan|other [seed] [iterations]      (test infrastructure; needs no GPU)
"""
import random
import sys
import time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
which = sys.argv[1] if len(sys.argv) > 1 else "scan"
sys.argv = [sys.argv[0]] + sys.argv[2:]
if which == 'scan':
    import torch, numpy as np
    import emu
    import test_gpu_scan as tg
    from util import rand_scan_inputs
    random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 0)
    fails=0
    t0=time.time()
    with emu.emulated():
        for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 40):
            batch=random.choice([1,1,2,3]); N=random.choice([8,16,16]); G=random.choice([1,1,1,2,3])
            dpg=random.choice([1,2,5,8,31,32,33,40,64,70]); dim=dpg*G
            L=random.choice([1,2,3,4,5,7,8,9,31,32,33,63,64,65,127,255,256,257,300,511,512,513,1000,2047,2048,2049,2600])
            dtype=random.choice([torch.float32,torch.float32,torch.bfloat16,torch.float16])
            direction=random.choice([0,1])
            has_D,has_z,has_b,sp=[random.random()<0.7 for _ in range(4)]
            use_h=random.random()<0.5
            d = rand_scan_inputs(1000+it, batch, dim, L, N, G, dtype, device="cpu", trained_like=sp)
            cfg=(batch,dim,L,N,G,str(dtype),direction,has_D,has_z,has_b,sp,use_h)
            try:
                res = tg._run_fwd_bwd(d, has_D, has_z, has_b, sp, direction=direction, use_hstates=use_h)
                ref = tg._oracle_fwd_bwd(d, has_D, has_z, has_b, sp, flip=bool(direction))
                tg._compare(res, ref, dtype, has_z)
            except Exception as e:
                fails+=1
                print("FAIL", cfg, type(e).__name__, str(e)[:200], flush=True)
    print("done fails=%d in %.1fs"%(fails, time.time()-t0))
    
else:
    import torch, numpy as np, torch.nn.functional as F
    import emu, golden_inputs as gi
    from util import rel_err
    from oracle import oracle as orc
    random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 0)
    fails=0; t0=time.time()
    def chk(name, a, b, tol, cfg):
        global fails
        try:
            e=rel_err(a,b)
            if not e<=tol: raise AssertionError("rel %.2e"%e)
        except Exception as ex:
            fails+=1; print("FAIL",name,cfg,str(ex)[:150],flush=True)
    with emu.emulated():
        from segmamba_b200 import causal_conv1d_cuda as cc
        from segmamba_b200.instance_norm import fused_instance_norm
        from segmamba_b200.layer_norm import fused_layer_norm, supported
        for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 30):
            # conv1d
            batch=random.choice([1,2,3]); dim=random.choice([1,3,8,31,32,33,96]); L=random.choice([1,2,3,4,7,8,9,255,256,257,1000,2049]); width=random.choice([2,3,4])
            dtype=random.choice([torch.float32,torch.bfloat16,torch.float16]); direction=random.choice([0,1]); silu=random.random()<0.7; hasb=random.random()<0.7
            d=gi.conv_inputs(500+it,batch,dim,L,width)
            x,dout,w,b=d["x"].to(dtype),d["dout"].to(dtype),d["weight"],(d["bias"] if hasb else None)
            cfg=("conv",batch,dim,L,width,str(dtype),direction,silu,hasb)
            try:
                out=cc.causal_conv1d_fwd_ex(x,w,b,silu,direction=direction)
                dx,dw,db=cc.causal_conv1d_bwd_ex(x,w,b,dout,None,silu,direction=direction)
                f=(lambda t:t.flip(-1)) if direction else (lambda t:t)
                o=f(orc.causal_conv1d_fwd_raw(f(x.float()),w,b,silu))
                odx,odw,odb=orc.causal_conv1d_bwd_raw(f(x.float()),w,b,f(dout.float()),silu)
                lo=dtype==torch.float32
                chk("out",out,o,1e-5 if lo else 1e-2,cfg); chk("dx",dx,f(odx),1e-4 if lo else 3e-2,cfg); chk("dw",dw,odw,1e-4 if lo else 3e-2,cfg)
                if hasb: chk("db",db,odb,1e-4 if lo else 3e-2,cfg)
            except Exception as ex:
                fails+=1; print("EXC",cfg,type(ex).__name__,str(ex)[:200],flush=True)
            # instnorm
            v=4 if dtype==torch.float32 else 8
            C=v*random.choice([1,2,3,6,12,24]); sp=(random.choice([1,2,3,5,8]),random.choice([1,2,4,7]),random.choice([1,3,4,9])); B=random.choice([1,2,3])
            mode=random.choice(["plain","add","addnorm"]); act=random.choice([None,"relu","leaky_relu"])
            cfg=("in",B,C,sp,str(dtype),mode,act)
            try:
                torch.manual_seed(it)
                xx=(torch.randn(B,C,*sp)*2+0.7).to(dtype).contiguous(memory_format=torch.channels_last_3d).requires_grad_()
                add=(torch.randn(B,C,*sp)*0.5-0.3).to(dtype).requires_grad_() if mode!="plain" else None
                dy=torch.randn(B,C,*sp).to(dtype)
                y=fused_instance_norm(xx,act,0.01,add=add,add_norm=(mode=="addnorm"))
                gx=torch.autograd.grad(y,[xx]+([add] if add is not None else []),dy)
                xr=xx.detach().float().requires_grad_(); ar=add.detach().float().requires_grad_() if add is not None else None
                vv=F.instance_norm(xr,eps=1e-5) if xr[0,0].numel()>1 else (xr-xr)   # torch refuses 1 spatial element
                if xr[0,0].numel()==1: raise RuntimeError("skip")
                if ar is not None: vv=vv+(F.instance_norm(ar,eps=1e-5) if mode=="addnorm" else ar)
                if act=="relu": vv=F.relu(vv)
                elif act=="leaky_relu": vv=F.leaky_relu(vv,0.01)
                gr=torch.autograd.grad(vv,[xr]+([ar] if ar is not None else []),dy.float())
                lo=dtype==torch.float32
                chk("y",y,vv,2e-4 if lo else 1e-2,cfg)
                for a_,b_ in zip(gx,gr): chk("dx",a_,b_,1e-3 if lo else 5e-2,cfg)
            except RuntimeError as ex:
                if "skip" not in str(ex): fails+=1; print("EXC",cfg,str(ex)[:200],flush=True)
            # layernorm
            rows=random.choice([1,2,7,31,32,33,100,1025]); C2=v*random.choice([1,2,4,6,12,13,24,48,96])
            cfg=("ln",rows,C2,str(dtype))
            xx=(torch.randn(rows,C2)*1.3+0.2).to(dtype).requires_grad_()
            if supported(xx,C2):
                w2=(torch.rand(C2)+0.5).requires_grad_(); b2=torch.randn(C2).requires_grad_(); dy=torch.randn(rows,C2).to(dtype)
                y=fused_layer_norm(xx,w2,b2,1e-5); g=torch.autograd.grad(y,[xx,w2,b2],dy)
                xr=xx.detach().float().requires_grad_(); wr=w2.detach().clone().requires_grad_(); br=b2.detach().clone().requires_grad_()
                yr=F.layer_norm(xr,(C2,),wr,br,1e-5); gr=torch.autograd.grad(yr,[xr,wr,br],dy.float())
                lo=dtype==torch.float32
                chk("y",y,yr,1e-5 if lo else 1e-2,cfg)
                for a_,b_ in zip(g,gr): chk("g",a_,b_,2e-4 if lo else 3e-2,cfg)
    print("done fails=%d in %.1fs"%(fails,time.time()-t0))
    
