import ctypes, torch, sys, numpy as np
sys.path.insert(0,'.')
from hvpr_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libhvpr_b200.so","libhvpr_b200_prof.so")
_lib.init_device(); L=_lib.lib()
from hvpr_b200 import synth as hybrid_syn
from hvpr_b200 import synth
from hvpr_b200.geometry import G2
from hvpr_b200.frontend import HybridFrontEnd
w = hybrid_syn.random_frontend_weights(0)
fe = HybridFrontEnd(G2, mem_precision="fp32").load_reference_weights(w)
B,N=8,120000
frames = synth.make_batch("L", N, G2.point_cloud_range, B)
p = fe.plan(B,B*N,N,use_graph=False)
p.points.copy_(torch.from_numpy(np.concatenate(frames,0))); p.frame_offsets.copy_(torch.tensor(np.r_[0,np.cumsum([N]*B)],dtype=torch.int32))
fe.run(); torch.cuda.synchronize()
P=int(p.vox.voxel_offsets[-1]); print("P",P)
pil = p.pillar_features[:P].contiguous(); Wd = fe.map_to_bev_module.memory.weight.detach()
wpk = torch.empty((2048,64),dtype=torch.bfloat16,device="cuda")
_lib.check(L.hvpr_mem_pack_bf16(_lib.ptr(Wd),2000,64,_lib.ptr(wpk),_lib.cur_stream()))
nb = L.hvpr_mem_attn_workspace_bytes(P,2000,1); ws=torch.empty(nb,dtype=torch.uint8,device="cuda")
out=torch.empty((P,64),device="cuda")
L.hvpr_dbg_mem_attn_logits.restype=ctypes.c_int
L.hvpr_dbg_mem_attn_logits.argtypes=[ctypes.c_void_p,ctypes.c_int64,ctypes.c_void_p,ctypes.c_void_p,ctypes.c_int,ctypes.c_void_p,ctypes.c_void_p,ctypes.c_void_p,ctypes.c_size_t,ctypes.c_void_p,ctypes.c_void_p]
for rep in range(2):
    prof=torch.zeros((148,24),dtype=torch.int64,device="cuda")
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.hvpr_dbg_mem_attn_logits(_lib.ptr(pil),P,_lib.ptr(Wd),_lib.ptr(wpk),2000,_lib.ptr(out),None,_lib.ptr(ws),nb,_lib.ptr(prof),_lib.cur_stream()))
    e1.record(); torch.cuda.synchronize()
    print("ms",e0.elapsed_time(e1))
pr=prof.cpu().numpy().astype(np.float64)
m=pr.mean(0)
names=["mma.wait_a_full","mma.wait_w_full","mma.wait_t_empty","flt.s1.sort+merge","flt.wait_tfull_s1","flt.wait_tfull_s2","flt.wait_c_empty","flt.s1.ld+max","tail.wait_c_full","tail.work","row.gather+dot","row.butterfly","row.select+softmax","row.readout+store","flt.s2.loop"]
g0,g1=pr[:,16],pr[:,17]
print(f"globaltimer: first start -> last end {(g1.max()-g0.min())/1e3:.1f} us; start spread {(g0.max()-g0.min())/1e3:.1f} us; end spread {(g1.max()-g1.min())/1e3:.1f} us; per-CTA span mean {(g1-g0).mean()/1e3:.1f} us")
mhz=pr[:,15]/((g1-g0)/1e3)
print("implied SM clock MHz: min %.0f mean %.0f max %.0f" % (mhz.min(),mhz.mean(),mhz.max()))
print("rows by tail path: fast %d medium %d slow %d; per-CTA slow rows max %d; per-CTA medium max %d" % (pr[:,19].sum(),pr[:,20].sum(),pr[:,21].sum(),pr[:,21].max(),pr[:,20].max()))
span=(g1-g0)/1e3
import numpy as _n
print("corr(span, slow rows) %.2f corr(span, medium rows) %.2f corr(span,rows) %.2f" % (_n.corrcoef(span,pr[:,21])[0,1], _n.corrcoef(span,pr[:,20])[0,1], _n.corrcoef(span,pr[:,19]+pr[:,20]+pr[:,21])[0,1]))
print("span percentiles us", _n.percentile(span,[0,10,50,90,100]).round(1), "rows/CTA", _n.percentile(pr[:,19]+pr[:,20]+pr[:,21],[0,50,100]))
o=_n.argsort(-span)[:10]
print("slowest CTAs (block, smid, span us, rows, medium rows, tail.work kcyc, select kcyc, gather kcyc):")
for b in o: print("  ", b, int(pr[b,18]), round(span[b],1), int(pr[b,19]+pr[b,20]), int(pr[b,20]), round(pr[b,9]/1e3), round(pr[b,12]/1e3), round(pr[b,10]/1e3))
o=_n.argsort(span)[:4]
for b in o: print(" fast", b, int(pr[b,18]), round(span[b],1), int(pr[b,19]+pr[b,20]), int(pr[b,20]), round(pr[b,9]/1e3), round(pr[b,12]/1e3), round(pr[b,10]/1e3))
print(f"CTA lifetime mean {m[15]/1e3:.1f} max {pr[:,15].max()/1e3:.1f} kcycles")
for i,n in enumerate(names): print(f"{n:20s} {m[i]/1e3:10.1f} kcycles")
