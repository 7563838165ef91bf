"""Re-export of musicstyletransfer_b200.MIDIUtil.defaults under the reference's module path."""
from musicstyletransfer_b200.MIDIUtil.defaults import *  # noqa: F401,F403
