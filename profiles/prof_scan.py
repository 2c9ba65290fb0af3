"""Small driver for the ncu capture of the fused scan kernel (single query and 8 queries, 1 M x 512, top-10).
    python profiles/prof_scan.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 4 -c 4 -o gpurun_out/prof_scan_r1c python profiles/prof_scan.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gofeature_b200 import api

ctx = api.NewContext(0)
cache = api.NewCache(ctx, 63, 32 << 20)
cache.NewSet("prof", 512, 4, 16)
st = cache.GetSet("prof")
st.AddSynthetic(1_000_000, 20260101)
rng = np.random.default_rng(0)
stream = torch.cuda.current_stream().cuda_stream
for q in (1, 1, 1, 1, 1, 1, 8, 8):
    x = rng.uniform(-1, 1, (q, 512)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    dq = torch.from_numpy(x).cuda()
    keys = torch.zeros((q, 10), dtype=torch.int64, device="cuda")
    st.SearchDevice(10, q, dq.data_ptr(), keys.data_ptr(), 1, stream)
    torch.cuda.synchronize()
print("ok", api.key_slot(int(keys.cpu().numpy().view(np.uint64)[0, 0])))
cache.Close()
ctx.Close()
