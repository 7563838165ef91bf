"""Re-export of musicstyletransfer_b200.VarAutoEncoder.loss under the reference's module path."""
from musicstyletransfer_b200.VarAutoEncoder.loss import *  # noqa: F401,F403
