"""Re-export of musicstyletransfer_b200.VarAutoEncoder.data under the reference's module path."""
from musicstyletransfer_b200.VarAutoEncoder.data import *  # noqa: F401,F403
