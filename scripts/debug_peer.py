import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ".")
import torch
import rxinfer_jl_b200 as rx
from rxinfer_jl_b200.sharding import PeerGroup
cs = [rx.Context(0, use_torch_stream=False) for _ in range(2)]
groups = PeerGroup.local(cs, 4, 1, 32, with_cov=False)
G, n = 2, 4000
bufs = [rx.context.DeviceBuffer(c, 4 * G * n) for c in cs]
loc = [torch.randn(n, device="cuda") for _ in cs]
ptrs = [bf.ptr for bf in bufs]
torch.cuda.synchronize()
t0 = time.perf_counter()
for r, c in enumerate(cs):
    c.peer_allgather(loc[r], ptrs, asynchronous=True)
    print("issued rank", r, "at", round(time.perf_counter() - t0, 4), flush=True)
for r, c in enumerate(cs):
    try:
        c.sync(); print("sync rank", r, "ok at", round(time.perf_counter() - t0, 4), flush=True)
    except Exception as e:
        print("sync rank", r, "ERR at", round(time.perf_counter() - t0, 4), e, flush=True)
print("flags0", groups[0].buf_flags.tensor(8, dtype=torch.int32).cpu().tolist(), "flags1", groups[1].buf_flags.tensor(8, dtype=torch.int32).cpu().tolist())
t = bufs[0].tensor(G, n)
print("slab0 ok", torch.equal(t[0], loc[0]), "slab1 ok", torch.equal(t[1], loc[1]))
# second attempt: rank 1 first
for r, c in reversed(list(enumerate(cs))):
    c.peer_allgather(loc[r], ptrs, asynchronous=True)
for r, c in enumerate(cs):
    try:
        c.sync(); print("2nd sync rank", r, "ok at", round(time.perf_counter() - t0, 4), flush=True)
    except Exception as e:
        print("2nd sync rank", r, "ERR", e, flush=True)
