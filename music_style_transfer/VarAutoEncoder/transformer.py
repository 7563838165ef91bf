"""Re-export of musicstyletransfer_b200.VarAutoEncoder.transformer under the reference's module path."""
from musicstyletransfer_b200.VarAutoEncoder.transformer import *  # noqa: F401,F403
