"""Re-export of musicstyletransfer_b200.VarAutoEncoder.trainer under the reference's module path."""
from musicstyletransfer_b200.VarAutoEncoder.trainer import *  # noqa: F401,F403
