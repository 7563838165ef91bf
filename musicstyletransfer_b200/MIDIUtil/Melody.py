"""Melody container and token events.

Public names, ids, ranges and error behaviour follow the reference's MIDIUtil/Melody.py (``Melody``, ``Event`` and
its three kinds, ``get_melody_from_ids``, ``create_event_from_id``, ``create_note_on_event``,
``create_note_off_event``, ``create_timeshift_event``) so that reader, writer, dataset and sampler code written
against the reference keeps working.  The implementation is table driven: every event kind is one row of
``_KINDS`` (class, first id, last id of its vocabulary range, MIDIUtil/defaults.py:43-58) and the id <-> event
conversions are range look-ups in that table.  ``get_midi_event`` builds events of the in-repo SMF module
(``MIDIUtil/smf.py``) because python-midi is not installable here.
"""
import bisect

from . import defaults as _d
from . import smf as midi
from .defaults import *  # noqa: F401,F403  (the reference re-exports the vocabulary constants from this module)
from .defaults import MAX_TICKS, MIN_TICKS, NUM_BINS, NUM_TICKS_IN_A_BIN  # noqa: F401


class Melody:
    """An ordered list of token events plus the header data a MIDI file needs (Melody.py:6-32)."""

    _META = ("key", "bpm", "resolution", "slices_per_quarter", "description")

    def __init__(self, key=_d.PITCH_C, bpm=_d.DEFAULT_BPM, resolution=_d.DEFAULT_RESOLUTION, slices_per_quarter=4,
                 description: str = ''):
        self.key, self.bpm, self.resolution = key, bpm, resolution
        self.slices_per_quarter = int(slices_per_quarter)
        self.description = description
        self.notes = []

    def __len__(self):
        return len(self.notes)

    def __getitem__(self, index):
        return self.notes[index]

    def __iter__(self):
        return iter(self.notes)

    def ids(self):
        """Token ids of the events, in order."""
        return [e.id for e in self.notes]

    def copy_metainformation(self):
        """A Melody with the same header data and no events."""
        return Melody(**{name: getattr(self, name) for name in self._META})


class Event:
    """A token of the event vocabulary; ``shifted_id`` is its index inside the range of its kind."""

    first_id = None                      # set per kind below

    def __init__(self, id):
        self.id = id

    @property
    def shifted_id(self):
        if self.first_id is None:
            raise NotImplementedError
        return int(self.id - self.first_id)

    def get_midi_event(self, tick_delay: int):
        raise NotImplementedError

    def __repr__(self):
        return "%s(%d)" % (type(self).__name__, self.id)


class NoteOnEvent(Event):
    first_id = _d.NOTE_ON_EVENTS[0]

    def get_midi_event(self, tick_delay: int):
        # the writer always plays notes at full velocity (Melody.py:56-58)
        return midi.NoteOnEvent(tick=tick_delay, pitch=self.shifted_id, velocity=127)


class NoteOffEvent(Event):
    first_id = _d.NOTE_OFF_EVENTS[0]

    def get_midi_event(self, tick_delay: int):
        return midi.NoteOffEvent(tick=tick_delay, pitch=self.shifted_id)


class TimeshiftEvent(Event):
    first_id = _d.TIMESHIFT_EVENTS[0]

    def get_tick_delay(self):
        """Ticks this token advances the clock by when a stream is played back (Melody.py:82-83)."""
        return self.shifted_id * NUM_TICKS_IN_A_BIN


# (class, first id, last id), ascending and contiguous: 3..130, 131..258, 259..292
_KINDS = sorted(((cls, rng[0], rng[1]) for cls, rng in ((NoteOnEvent, _d.NOTE_ON_EVENTS), (NoteOffEvent, _d.NOTE_OFF_EVENTS),
                                                       (TimeshiftEvent, _d.TIMESHIFT_EVENTS))), key=lambda k: k[1])
_FIRST_IDS = [k[1] for k in _KINDS]
_LOWEST, _END = _KINDS[0][1], _d.NUM_EVENTS


def create_event_from_id(id):
    """id -> event of the kind whose range holds it; ValueError outside [first NOTE_ON id, NUM_EVENTS)."""
    if not _LOWEST <= id < _END:
        raise ValueError("ID {} is not in range [{}, {}]".format(id, _LOWEST, _END))
    return _KINDS[bisect.bisect_right(_FIRST_IDS, id) - 1][0](id)


def get_melody_from_ids(ids):
    """Token ids -> Melody; PAD / SOS / EOS (ids below FEATURE_OFFSET) are dropped."""
    melody = Melody()
    melody.notes = [create_event_from_id(int(i)) for i in ids if i >= _d.FEATURE_OFFSET]
    return melody


def create_note_on_event(pitch: int):
    return NoteOnEvent(NoteOnEvent.first_id + pitch)


def create_note_off_event(pitch: int):
    return NoteOffEvent(NoteOffEvent.first_id + pitch)


def create_timeshift_event(timeshift_ticks: int):
    """Ticks -> the time-shift token of the 30-tick bin that holds them (Melody.py:117-126); asserts the range."""
    assert MIN_TICKS <= timeshift_ticks < MAX_TICKS, \
        "Time shift must be between {} ticks and {} ticks. It is {}.".format(MIN_TICKS, MAX_TICKS, timeshift_ticks)
    event_id = TimeshiftEvent.first_id + int((timeshift_ticks - MIN_TICKS) / NUM_TICKS_IN_A_BIN)
    assert event_id <= _d.TIMESHIFT_EVENTS[1]
    return TimeshiftEvent(event_id)
