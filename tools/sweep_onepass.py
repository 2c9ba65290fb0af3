"""Small-batch sweep of gf_set_search_device_ex: the one-pass chain (path 2) against the two-pass chain (path 4) on
the same set, per shadow format; checks that both chains return identical keys.  N= / SHADOWS= in the environment."""
import sys; sys.path.insert(0,'.')
import numpy as np, torch, time, os
from gofeature_b200 import api
from oracle import lib as ol
ctx=api.NewContext(0)
L=api.lib()
N=int(os.environ.get("N","1000000"))
shadows=os.environ.get("SHADOWS","int8,bf16,none").split(",")
for shadow in shadows:
    cache=api.NewCache(ctx,(N+16383)//16384+1,32<<20,shadow=False if shadow=="none" else shadow); cache.NewSet("d",512,4,64); st=cache.GetSet("d")
    st.AddSynthetic(N,20260101)
    for Q,K in ((1,10),(10,10),(16,10),(32,10),(10,100)):
        q=ol.synth_rows(555+Q,0,Q,512)
        dq=torch.from_numpy(q).cuda(); s=torch.cuda.current_stream().cuda_stream
        res={}
        for path in (4,2):
            k2=torch.zeros((Q,K),dtype=torch.int64,device="cuda"); fl=torch.zeros((Q,),dtype=torch.int32,device="cuda")
            def run():
                api._check(L.gf_set_search_device_ex(st._h,K,Q,dq.data_ptr(),k2.data_ptr(),fl.data_ptr(),path,s))
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
            n=50
            for _ in range(n): run()
            e1.record(); torch.cuda.synchronize(); t_graph=e0.elapsed_time(e1)/n
            ctx.ResetStats(); ctx.SetProfiling(True)
            for _ in range(n): run()
            torch.cuda.synchronize(); stt=ctx.Stats(); ctx.SetProfiling(False)
            gm=stt['gemm_kernel_ns']/1e6/max(stt['gemm_kernel_timed'],1)
            res[path]=(k2.cpu().numpy().copy(), fl.cpu().numpy().copy())
            print("shadow=%s Q=%3d K=%3d path=%d step %.4f ms (%.0f qps) big kernel avg %.4f ms x%d flagged %d"%(shadow,Q,K,path,t_graph,Q/t_graph*1e3,gm,stt['gemm_kernel_timed']//n,int(res[path][1].sum())),flush=True)
        same=(res[2][0]==res[4][0]).all()
        print("   keys identical between chains:", bool(same), flush=True)
    cache.Close()
