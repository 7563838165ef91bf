"""Re-export of musicstyletransfer_b200.VarAutoEncoder.config under the reference's module path."""
from musicstyletransfer_b200.VarAutoEncoder.config import *  # noqa: F401,F403
