"""Re-export of musicstyletransfer_b200.VarAutoEncoder.utils under the reference's module path."""
from musicstyletransfer_b200.VarAutoEncoder.utils import *  # noqa: F401,F403
