"""int8 COLLECT pass at large batches (flag-less gf_set_search_device, 1 M x 512): step time and the pass kernel's own
time (CUDA events inside the library), for A/B runs of the epilogue."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from gofeature_b200 import api
from oracle import lib as ol
ctx = api.NewContext(0)
N = 1000000
cache = api.NewCache(ctx, (N + 16383) // 16384 + 1, 32 << 20, shadow="int8"); cache.NewSet("d", 512, 4, 1024); st = cache.GetSet("d")
st.AddSynthetic(N, 20260101)
for Q, K in ((64, 10), (128, 10), (256, 10), (256, 100), (1024, 100)):
    q = ol.synth_rows(555 + Q, 0, Q, 512)
    dq = torch.from_numpy(q).cuda(); s = torch.cuda.current_stream().cuda_stream
    k2 = torch.zeros((Q, K), dtype=torch.int64, device="cuda")
    for _ in range(3): st.SearchDevice(K, Q, dq.data_ptr(), k2.data_ptr(), 2, s)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    n = 30
    for _ in range(n): st.SearchDevice(K, Q, dq.data_ptr(), k2.data_ptr(), 2, s)
    e1.record(); torch.cuda.synchronize(); t_step = e0.elapsed_time(e1) / n
    ctx.ResetStats(); ctx.SetProfiling(True)
    for _ in range(n): st.SearchDevice(K, Q, dq.data_ptr(), k2.data_ptr(), 2, s)
    torch.cuda.synchronize(); stt = ctx.Stats(); ctx.SetProfiling(False)
    gm = stt['gemm_kernel_ns'] / 1e6 / max(stt['gemm_kernel_timed'], 1)
    print("Q=%4d K=%3d step %.4f ms (%.0f qps)  collect avg %.4f ms x%d -> %.0f GB/s, %.2f POP/s useful" % (
        Q, K, t_step, Q / t_step * 1e3, gm, stt['gemm_kernel_timed'] // n, N * 512 / gm / 1e6, 2 * min(Q, 256) * N * 512 / gm / 1e12), flush=True)
cache.Close()
