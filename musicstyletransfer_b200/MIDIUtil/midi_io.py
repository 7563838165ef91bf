"""MIDI reader / writer with the reference's interface (MIDIUtil/midi_io.py).

``EventBasedMIDIReader.read_file`` keeps its contract (list of Melody, tracks with < 10 tokens dropped,
at least one must survive) but the per-event featurisation loop of ``_parse_track`` (midi_io.py:70-93)
runs on the GPU: the C++ parser in libmsx.so (``msx_smf_parse``) turns the .mid bytes of every track into the
note-event SoA (dtick since the previous note event, pitch, velocity) and ``featurise.tokenize_tracks`` launches the
K1 kernel once for all tracks of all files handed to ``read_files``.  The pure-Python SMF module (``smf.py``) remains for
writing files (``MelodyWriter``) and as an independent reader in the tests."""
import numpy as np

from . import smf as midi
from .Melody import *  # noqa: F401,F403
from .Melody import Melody, NoteOffEvent, NoteOnEvent, TimeshiftEvent, create_event_from_id
from .defaults import *  # noqa: F401,F403
from .defaults import DEFAULT_BPM, N_PITCHES


class MIDIReader():
    def __init__(self, slices_per_quarter_note: int):
        self.slices_per_quarter_note = slices_per_quarter_note
        print("Time slices per quarter note: {}".format(self.slices_per_quarter_note))

    def _extract_bpm(self, pattern):
        for track in pattern:
            for event in track:
                if isinstance(event, midi.SetTempoEvent):
                    return event.get_bpm()
        return DEFAULT_BPM

    def read_file(self, file_name):
        raise NotImplementedError


def note_event_soa(track):
    """Walk a parsed track as midi_io.py:73-91 does: every event's tick advances the clock (:75), only
    note events reset ``prev_t`` (:91)."""
    dt, pi, ve = [], [], []
    prev_t = cur_t = 0
    for ev in track:
        cur_t += ev.tick
        if isinstance(ev, (midi.NoteOnEvent, midi.NoteOffEvent)):
            dt.append(cur_t - prev_t)
            pi.append(ev.data[0])
            ve.append(ev.data[1])
            prev_t = cur_t
    return (np.asarray(dt, dtype=np.int32), np.asarray(pi, dtype=np.uint8), np.asarray(ve, dtype=np.uint8))


class EventBasedMIDIReader(MIDIReader):
    def __init__(self):
        super().__init__(0)

    def read_files(self, file_names):
        """Batched form: one kernel launch tokenises every track of every file.  Returns {file: [Melody]}."""
        from .. import featurise
        # .mid bytes -> note-event SoA in the C++ parser of libmsx.so (msx_smf_parse), SoA -> token ids in K1
        parsed = [featurise.parse_smf_file(f) for f in file_names]
        soas, owners = [], []
        for fi, (_, tracks, _) in enumerate(parsed):
            for soa in tracks:
                soas.append(soa)
                owners.append(fi)
        ids_per_track = featurise.tokenize_tracks(soas)
        out = {}
        for fi, (fname, (info, _, _)) in enumerate(zip(file_names, parsed)):
            result = []
            for ti, owner in enumerate(owners):
                if owner != fi:
                    continue
                new_melody = Melody(bpm=info["bpm"], resolution=info["resolution"], slices_per_quarter=self.slices_per_quarter_note)
                new_melody.notes = [create_event_from_id(int(i)) for i in ids_per_track[ti]]
                new_melody.soa = soas[ti]                        # note-event SoA: lets the device dataset re-tokenise in place
                if len(new_melody) < 10:                         # midi_io.py:60-63
                    print('Warning: {} contains melodies of length {} < 10. Discarding'.format(fname, len(new_melody.notes)))
                    continue
                result.append(new_melody)
            assert len(result) > 0                               # midi_io.py:67
            out[fname] = result
        return out

    def read_file(self, file_name):
        return self.read_files([file_name])[file_name]


class MelodyWriter:
    def __init__(self):
        self.tempo = DEFAULT_BPM

    def get_midi_pitch(self, note):
        return note.octave * N_PITCHES + note.pitch

    def write_to_file(self, file_name, melody):
        pattern = midi.Pattern()
        pattern.resolution = melody.resolution
        track = midi.Track()
        track.append(self._create_bpm_event(melody))
        self._write_track(melody, track)
        track.append(midi.EndOfTrackEvent(tick=1))
        pattern.append(track)
        midi.write_midifile(file_name, pattern)

    def _write_track(self, melody, track):
        tick_delay = 0
        for event in melody:
            if isinstance(event, TimeshiftEvent):
                tick_delay += event.get_tick_delay()
            elif isinstance(event, (NoteOnEvent, NoteOffEvent)):
                track.append(event.get_midi_event(int(tick_delay)))
                tick_delay = 0

    def _create_bpm_event(self, melody):
        bpm_event = midi.SetTempoEvent()
        bpm_event.set_bpm(melody.bpm)
        return bpm_event
