"""Re-export of musicstyletransfer_b200.VarAutoEncoder.sampler under the reference's module path (CLI entry point)."""
from musicstyletransfer_b200.VarAutoEncoder.sampler import *  # noqa: F401,F403
from musicstyletransfer_b200.VarAutoEncoder.sampler import main

if __name__ == "__main__":
    main()
