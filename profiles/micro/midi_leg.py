import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import bench, torch
from musicstyletransfer_b200.engine import VAEConfig
sys.argv = ["bench.py"]
args = bench.parse_args()
cfg = VAEConfig(dec_type="lstm", enc_dropout=args.dropout, dec_dropout=args.dropout)
for i in range(3):
    r = bench.bench_from_midi(args, cfg, "cuda:0")
    print(i, round(r["value"]), {k: round(v, 1) for k, v in r["stage_ms"].items()})
