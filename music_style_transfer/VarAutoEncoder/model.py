"""Re-export of musicstyletransfer_b200.VarAutoEncoder.model under the reference's module path."""
from musicstyletransfer_b200.VarAutoEncoder.model import *  # noqa: F401,F403
