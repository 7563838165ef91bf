"""Per-entry-point time table of one style-transfer pass (config 5: B = 8192, L = 64, LSTM decoder, 130 sampled steps)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from musicstyletransfer_b200 import lib, synth
from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
dev = torch.device("cuda", 0)
eng = VAEEngine(VAEConfig(dec_type="lstm"), dev, seed=0, precision="tf32")
B, L = 8192, 64
tok, lens, cls, lab = synth.token_rows_4_4(B, L, seed=11)
t = lambda a: torch.from_numpy(a).to(dev)
tokens, seq_lens = t(tok), t(lens)
target = torch.zeros(B, dtype=torch.int32, device=dev)
eng.style_transfer(tokens, seq_lens, target, seed=1)
torch.cuda.synchronize()
kp = {"match": "msx_", "events": [], "names": True}
lib._profile = kp
eng.style_transfer(tokens, seq_lens, target, seed=2)
torch.cuda.synchronize()
lib._profile = None
tab = {}
for a, b, f in kp["events"]:
    name = f if isinstance(f, str) else f[2].split(" sk=")[0]
    e = tab.setdefault(name, [0, 0.0]); e[0] += 1; e[1] += a.elapsed_time(b)
tot = sum(v[1] for v in tab.values())
for k, v in sorted(tab.items(), key=lambda kv: -kv[1][1]):
    print("%-56s n=%4d %8.1f us/call %8.2f ms %5.1f%%" % (k, v[0], 1e3 * v[1] / v[0], v[1], 100 * v[1] / tot))
