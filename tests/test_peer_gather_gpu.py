"""Fused sweep + all-gather over peer-mapped buffers (rxg_peer.cu): every rank's gathered buffers must hold exactly
what independent per-shard sweeps produce, for the fused-store path (shared model), the push path (masks, per-chain
path, large-state family) and both covariance variants (full gather / RXG_COV_REPLICATE: bit-identical).
Ranks are contexts with their own streams inside one process on one GPU (no IPC), plus a two-PROCESS run that goes
through the CUDA IPC handles exactly as a multi-GPU job does (both processes may share cuda:0)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import lgssm
from util import TOL_COV, TOL_MEAN, f32_model, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _quiet_host:
    """Virtual ranks share ONE process here: while rank 0's barrier kernel spins on the device waiting for rank 1, the host
    must not run anything that synchronises the whole device before rank 1's work is queued -- in particular the cyclic
    garbage collector freeing DeviceBuffers of an earlier test (cudaFree waits for every running kernel).  Real jobs run
    one process per GPU and are not exposed to this."""

    def __enter__(self):
        import gc
        gc.collect()
        torch.cuda.synchronize()
        gc.disable()

    def __exit__(self, *a):
        import gc
        gc.enable()


def _kw(mod):
    return dict(A=mod["A"], B=mod["B"], P=mod["P"], Q=mod["Q"], m0=mod["m0"], S0=mod["S0"])


@pytest.fixture(scope="module")
def ranks(rx):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cs = [rx.Context(0, use_torch_stream=False) for _ in range(3)]
    yield cs
    for c in cs:
        c.close()


def test_two_processes_cuda_ipc():
    """The real multi-process protocol: handles exported, exchanged over torch.distributed (gloo), opened with
    cudaIpcOpenMemHandle; the two ranks use cuda:0 and cuda:1 when there are two GPUs, else share cuda:0."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "workers", "peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("PEER_WORKER_OK") == 2


@pytest.mark.parametrize("G", [2, 3])
@pytest.mark.parametrize("variant", ["full", "replicate", "masked", "per_chain_path", "full_push", "replicate_push"])
def test_virtual_ranks_fused_gather(rx, ranks, G, variant):
    from rxinfer_jl_b200.sharding import PeerGroup
    mod = f32_model(lgssm.notebook_model(4))
    T, b = 60, 96
    _, y = lgssm.generate_data(mod, T, G * b, seed=101)
    rng = np.random.default_rng(5)
    mask = (rng.random((T, G * b)) > 0.25).astype(np.uint8) if variant == "masked" else None
    cs = ranks[:G]
    push = variant.endswith("_push")                        # RXG_OPT_GATHER_MODE = 2: plain sweep, then peer_push_kernel
    variant = variant[:-5] if push else variant
    for c in cs:                                            # (module-scoped contexts: set it every time)
        c.set_option("gather_mode", 2 if push else 0)
    groups = PeerGroup.local(cs, T, 4, b)
    kw = dict(replicate_cov=(variant == "replicate"), force_per_chain_path=(variant == "per_chain_path"))
    refs = []
    for r, c in enumerate(cs):                              # reference: the plain sweep per shard (also sizes the workspace)
        ys = torch.as_tensor(np.ascontiguousarray(y[:, :, r * b:(r + 1) * b]), device="cuda")
        ms = torch.as_tensor(np.ascontiguousarray(mask[:, r * b:(r + 1) * b]), device="cuda") if mask is not None else None
        refs.append((ys, ms, c.lgssm(ys, **_kw(mod), smooth=True, mask=ms, force_per_chain_path=kw["force_per_chain_path"])))
    for rep in range(2):                                    # twice: the barrier epochs advance
        for gr in groups:
            gr.mean.zero_(); gr.cov.zero_()
        torch.cuda.synchronize()
        with _quiet_host():
            for r, gr in enumerate(groups):
                gr.smooth_gather(refs[r][0], mod, mask=refs[r][1], asynchronous=True, **kw)
            for c in cs:
                c.sync()
        for gr in groups:
            for r in range(G):
                assert torch.equal(gr.mean[r], refs[r][2]["mean"]), (variant, rep, gr.rank, r)
                assert torch.equal(gr.cov[r], refs[r][2]["cov"]), (variant, rep, gr.rank, r)
    # and against the oracle once, through the assembled layout
    full = rx.sharding.assemble_gathered(groups[0].mean).cpu().numpy()
    ref = lgssm.smooth_reference_schedule(y, **mod, mask=mask)
    assert rel_l2(full, ref["mean"]) < TOL_MEAN
    assert rel_l2(rx.sharding.assemble_gathered(groups[0].cov).cpu().numpy(), ref["cov"]) < TOL_COV


def test_virtual_ranks_large_state(rx, ranks):
    """d = 16 (tensor-core family: no in-kernel peer stores => slab push)."""
    from rxinfer_jl_b200.sharding import PeerGroup
    mod = f32_model(lgssm.dense_model(16))
    T, b, G = 30, 64, 2
    _, y = lgssm.generate_data(mod, T, G * b, seed=7)
    cs = ranks[:G]
    groups = PeerGroup.local(cs, T, 16, b)
    refs = []
    for r, c in enumerate(cs):
        ys = torch.as_tensor(np.ascontiguousarray(y[:, :, r * b:(r + 1) * b]), device="cuda")
        refs.append((ys, c.lgssm(ys, **_kw(mod), smooth=True)))
    for replicate in (False, True):
        for gr in groups:
            gr.mean.zero_(); gr.cov.zero_()
        torch.cuda.synchronize()
        with _quiet_host():
            for r, gr in enumerate(groups):
                gr.smooth_gather(refs[r][0], mod, replicate_cov=replicate, asynchronous=True)
            for c in cs:
                c.sync()
        for gr in groups:
            for r in range(G):
                assert torch.equal(gr.mean[r], refs[r][1]["mean"]) and torch.equal(gr.cov[r], refs[r][1]["cov"])


def test_virtual_ranks_generic_allgather(rx, ranks):
    """rxg_peer_allgather_f32 on arbitrary arrays: a 16-byte aligned slab (float4 path) and an odd-sized one (scalar path)."""
    G = 2
    cs = ranks[:G]
    groups = rx.sharding.PeerGroup.local(cs, 4, 1, 32, with_cov=False)   # registers the group; keeps the flag buffers alive
    assert len(groups) == G
    for n in (4 * 1000, 4 * 1000 + 3):
        bufs = [rx.context.DeviceBuffer(c, 4 * G * n) for c in cs]
        loc = [torch.randn(n, device="cuda") for _ in cs]
        ptrs = [bf.ptr for bf in bufs]
        with _quiet_host():
            for r, c in enumerate(cs):
                c.peer_allgather(loc[r], ptrs, asynchronous=True)
            for c in cs:
                c.sync()
        for bf in bufs:
            t = bf.tensor(G, n)
            for r in range(G):
                assert torch.equal(t[r], loc[r])
