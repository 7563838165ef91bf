"""Re-export of musicstyletransfer_b200.VarAutoEncoder.main under the reference's module path (CLI entry point)."""
from musicstyletransfer_b200.VarAutoEncoder.main import *  # noqa: F401,F403
from musicstyletransfer_b200.VarAutoEncoder.main import main

if __name__ == "__main__":
    main()
