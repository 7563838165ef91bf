"""Re-export of musicstyletransfer_b200.MIDIUtil.Melody under the reference's module path."""
from musicstyletransfer_b200.MIDIUtil.Melody import *  # noqa: F401,F403
