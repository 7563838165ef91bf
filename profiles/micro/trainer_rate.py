"""Trainer._step rate at the reference's own batch size (scripts/train-vae.sh: B = 32, L = 64) through the drop-in Trainer
class, eager launches vs CUDA-graph replay (synthetic 4/4 rows, train-vae.sh model, dropout 0.2)."""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from musicstyletransfer_b200 import synth
from musicstyletransfer_b200.VarAutoEncoder import main as vmain, model, trainer
from musicstyletransfer_b200.VarAutoEncoder.data import DataBatch

argv = ("--batch-size 32 --kl-loss 1.0 --max-seq-len 64 --e-n-layers 2 --e-dropout 0.2 --e-rnn-hidden-dim 256 "
        "--e-emb-hidden-dim 256 --latent-dim 256 --d-n-layers 1 --d-rnn-hidden-dim 128 --d-dropout 0.2 --optimizer adam "
        "--optimizer-params clip_gradient:1.0 --learning-rate 0.0003").split()
from musicstyletransfer_b200.VarAutoEncoder import config as cfgmod
args = cfgmod.get_config(argv)
for B in (32, 256):
    tok, lens, cls, lab = synth.token_rows_4_4(B * 8, 64, seed=1)
    batches = [DataBatch([torch.from_numpy(tok[i * B:(i + 1) * B].astype('float32')), torch.from_numpy(lens[i * B:(i + 1) * B].astype('float32')),
                          torch.from_numpy(cls[i * B:(i + 1) * B].astype('float32'))], [torch.from_numpy(lab[i * B:(i + 1) * B].astype('float32'))])
               for i in range(8)]
    for graph in (False, True):
        class _DS:                                  # the two dataset properties create_model_config reads
            def num_tokens(self): return 293
            def num_classes(self): return 2
        m = model.Model(vmain.create_model_config(args, _DS()), precision="tf32", quiet=True)
        t = trainer.Trainer(vmain.create_train_config(args), None, m, None, log_dir="/tmp/tb_rate", cuda_graph=graph)
        for i in range(20):
            t._step(batches[i % 8])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 300
        for i in range(n):
            t._step(batches[i % 8])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("B=%d graph=%s: %.3f ms/step, %.0f sequences/s (wall clock, host loop included)" % (B, graph, 1e3 * dt / n, B * n / dt))
