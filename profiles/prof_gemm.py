"""Small driver for the ncu capture of the tcgen05 COLLECT pass at a large batch (256 queries, 1 M x 512, top-10).
    python profiles/prof_gemm.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 6 -c 2 -o gpurun_out/prof_gemm256 python profiles/prof_gemm.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gofeature_b200 import api

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = api.NewContext(0)
cache = api.NewCache(ctx, 63, 32 << 20)
cache.NewSet("prof", 512, 4, 1024)
st = cache.GetSet("prof")
st.AddSynthetic(1_000_000, 20260101)
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, (Q, 512)).astype(np.float32)
x /= np.linalg.norm(x, axis=1, keepdims=True)
dq = torch.from_numpy(x).cuda()
keys = torch.zeros((Q, K), dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
ctx.SetProfiling(True)            # kernel by kernel (no graph replay) so that ncu sees plain launches
for _ in range(5):
    st.SearchDevice(K, Q, dq.data_ptr(), keys.data_ptr(), 2, stream)
    torch.cuda.synchronize()
print("ok", ctx.Stats()["tensor_passes"])
cache.Close()
ctx.Close()
