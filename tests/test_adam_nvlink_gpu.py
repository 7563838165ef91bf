"""Fused reduce-scatter + Adam + all-gather kernel (adam_nvlink.cu).  On one GPU the protocol is exercised with
VIRTUAL ranks: `world` independent arena sets on the same device, one stream per rank, so that the in-kernel flag
barriers, the slice ownership and the peer pushes run exactly as they do across NVLink; the result must equal the
baseline realisation of the trainer.py:176-177 rule (sum of the gradients in rank order, then msx_adam_step)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n", [(1, 4096), (2, 100000), (4, 2060328), (8, 40004)])
def test_virtual_ranks_match_allreduce_plus_adam(world, n):
    from musicstyletransfer_b200 import ops
    dev = "cuda"
    gen = torch.Generator().manual_seed(world * 7 + 1)
    w0 = torch.randn(n, generator=gen).to(dev)
    grads = [[torch.randn(n, generator=gen).to(dev) * (1.0 + r) for r in range(world)] for _ in range(3)]
    hyper = dict(lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0, rescale=1.0 / 64, clip=1.0)

    # baseline: all-reduce (sum in rank order) + msx_adam_step
    wr, mr, vr = w0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    state_r = torch.zeros(4, device=dev)
    for step in range(3):
        gs = grads[step][0].clone()
        for r in range(1, world):
            gs = gs + grads[step][r]
        ops.adam_step(wr, gs, mr, vr, n, state_r, hyper["lr"], hyper["beta1"], hyper["beta2"], hyper["eps"], hyper["wd"],
                      hyper["rescale"], hyper["clip"], zero_grad=True)

    # virtual ranks
    W = [w0.clone() for _ in range(world)]
    G = [torch.zeros(n, device=dev) for _ in range(world)]
    M = [torch.zeros(n, device=dev) for _ in range(world)]
    V = [torch.zeros(n, device=dev) for _ in range(world)]
    S = [torch.zeros(4, device=dev) for _ in range(world)]
    F = [torch.zeros(ops.adam_nvlink_flag_bytes() // 8, dtype=torch.int64, device=dev) for _ in range(world)]
    DC = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world)]
    EC = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for step in range(3):
        for r in range(world):
            G[r].copy_(grads[step][r])
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                # 12 CTAs per rank keeps all ranks co-resident on one GPU (they spin on each other's flags)
                ops.adam_nvlink_step(W[r], G[r], M[r], V[r], n, S[r], [t.data_ptr() for t in G], [t.data_ptr() for t in W],
                                     [t.data_ptr() for t in F], DC[r], r, world, EC[r], hyper["lr"], hyper["beta1"],
                                     hyper["beta2"], hyper["eps"], hyper["wd"], hyper["rescale"], hyper["clip"],
                                     zero_grad=True, max_ctas=12)
        torch.cuda.synchronize()
        for r in range(world):
            assert float(G[r].abs().max()) == 0.0          # zeroed for the next backward
    assert all(int(e.item()) == 3 for e in EC)
    for r in range(world):
        assert torch.equal(W[r], W[0]), "rank %d holds different parameters" % r
    assert float((W[0] - wr).abs().max()) <= 1e-7 * float(wr.abs().max())
    # optimiser state is sharded: rank r only ever touched the moments of its own slice
    n4 = n // 4
    per = (n4 + world - 1) // world
    for r in range(world):
        lo, hi = min(per * r, n4) * 4, min(per * (r + 1), n4) * 4
        assert torch.allclose(M[r][lo:hi], mr[lo:hi], rtol=1e-6, atol=1e-9)
        untouched = torch.cat([M[r][:lo], M[r][hi:]])
        assert untouched.numel() == 0 or float(untouched.abs().max()) == 0.0


def test_peer_optimizer_two_gpus():
    """Real NVLink path (needs >= 2 GPUs; the single-GPU boxes skip it): tests/dist_peer_check.py under torchrun."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(repo, "tests", "dist_peer_check.py")],
                         capture_output=True, text=True, timeout=300)
    assert "PEER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
