"""Re-export of musicstyletransfer_b200.MIDIUtil.midi_io under the reference's module path."""
from musicstyletransfer_b200.MIDIUtil.midi_io import *  # noqa: F401,F403
