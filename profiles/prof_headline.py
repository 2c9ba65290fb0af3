"""Small driver for the ncu captures of the headline step (1 M x 512, 10 queries top-10, int8 one-pass chain):
    python profiles/prof_headline.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k 'regex:gemm_kernel|grouptop_tail_kernel|prep_queries' -s 15 -c 6 \
        -f -o gpurun_out/prof_headline python profiles/prof_headline.py
bench.py repeats its steps for >= 1.2 s of device time, far too many launches for a profiler pass; this runs 16."""
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
rng = np.random.default_rng(20260102)
x = rng.uniform(-1, 1, (10, 512)).astype(np.float32)
x /= np.linalg.norm(x, axis=1, keepdims=True)
dq = torch.from_numpy(x).cuda()
keys = torch.zeros((10, 10), dtype=torch.int64, device="cuda")
flags = torch.zeros((10,), dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for _ in range(16):
    st.SearchDeviceEx(10, 10, dq.data_ptr(), keys.data_ptr(), flags.data_ptr(), 0, stream)
torch.cuda.synchronize()
print("ok", int(flags.sum().item()), api.key_slot(int(keys.cpu().numpy().view(np.uint64)[0, 0])))
cache.Close()
ctx.Close()
