"""Re-export of musicstyletransfer_b200.VarAutoEncoder.metrics under the reference's module path."""
from musicstyletransfer_b200.VarAutoEncoder.metrics import *  # noqa: F401,F403
