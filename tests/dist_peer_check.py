"""Run under torchrun on >= 2 GPUs: the fused NVLink optimiser step (reduce-scatter + Adam + all-gather over torch
symmetric memory) against the NCCL all-reduce + full-Adam baseline on the real train step, eager and CUDA-graph replay."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from musicstyletransfer_b200 import lib, synth  # noqa: E402
from musicstyletransfer_b200.engine import VAEConfig, VAEEngine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib.load()
    B, L = 48, 20
    tok, lens, cls, lab = synth.token_rows_4_4(B, L, seed=50 + rank)
    batch = [torch.from_numpy(a).to(dev) for a in (tok, lens, cls, lab)]
    cfg = VAEConfig(dec_type="lstm", enc_dropout=0.0, dec_dropout=0.0)
    results = {}
    for mode in ("nccl", "peer", "peer_graph"):
        eng = VAEEngine(cfg, dev, seed=3, precision="tf32")
        if mode == "nccl":
            ar = lambda g: dist.all_reduce(g)
        else:
            assert eng.enable_peer_optimizer() == world
            ar = "peer"
        step = eng.train_step_graphed if mode == "peer_graph" else eng.train_step
        for _ in range(4):
            step(*batch, kl_weight=1.0, global_batch=B * world, lr=3e-4, clip_gradient=1.0, allreduce=ar)
        torch.cuda.synchronize()
        w = eng.arena.w.clone()
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        same = all(torch.equal(ws[0], x) for x in ws)
        results[mode] = (w, same)
        eng._graphs.clear()
    ok = True
    for mode, (w, same) in results.items():
        d = float((w - results["nccl"][0]).abs().max())
        if rank == 0:
            print("%s: identical on all ranks = %s, max |w - w_nccl| = %.3e" % (mode, same, d), flush=True)
        # eps comes from the same seed sequence in all three runs; differences are atomic-order noise amplified by Adam
        ok = ok and same and d < 5e-3
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PEER_OK" if int(t.item()) == 1 else "PEER_FAIL", flush=True)
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
