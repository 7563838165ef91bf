"""Minimal Standard-MIDI-File reader for the oracle (test infrastructure).

Replaces the un-vendored third-party dependency ``midi`` (python-midi, un-pinned,
imported at MIDIUtil/midi_io.py:4 and MIDIUtil/Melody.py:1) for exactly the
surface the reference touches: ``midi.read_midifile`` (midi_io.py:39),
``pattern.resolution`` (:44), iteration over tracks/events (:51,:74),
``event.tick`` (:75), ``event.data`` (:80), the ``NoteOnEvent`` / ``NoteOffEvent``
/ ``SetTempoEvent`` classes (:23,:79) and ``SetTempoEvent.get_bpm`` (:24).

Published SMF 1.0 format: ``MThd`` header (format, ntrks, division) followed by
``MTrk`` chunks of <variable-length delta><event>, with running status for
channel messages, ``FF type len data`` meta events and ``F0/F7 len data`` sysex.
"""
import struct


class Event:
    name = "Event"

    def __init__(self, tick=0, data=None, channel=0):
        self.tick = tick
        self.data = list(data) if data is not None else []
        self.channel = channel

    def __repr__(self):
        return "%s(tick=%d, data=%r)" % (type(self).__name__, self.tick, self.data)


class NoteOnEvent(Event):
    pass


class NoteOffEvent(Event):
    pass


class OtherChannelEvent(Event):
    pass


class MetaEvent(Event):
    metacommand = None


class SetTempoEvent(MetaEvent):
    def get_mpqn(self):
        return (self.data[0] << 16) | (self.data[1] << 8) | self.data[2]

    def get_bpm(self):
        # python-midi: float(6e7) / mpqn
        return float(6e7) / self.get_mpqn()


class EndOfTrackEvent(MetaEvent):
    pass


class SysexEvent(Event):
    pass


class Track(list):
    pass


class Pattern(list):
    def __init__(self, resolution=220, format=1):
        super().__init__()
        self.resolution = resolution
        self.format = format


def _read_varlen(buf, pos):
    value = 0
    while True:
        b = buf[pos]
        pos += 1
        value = (value << 7) | (b & 0x7F)
        if not b & 0x80:
            return value, pos


# number of data bytes per channel-message status nibble
_CHANNEL_LEN = {0x8: 2, 0x9: 2, 0xA: 2, 0xB: 2, 0xC: 1, 0xD: 1, 0xE: 2}


def parse_bytes(buf):
    if buf[:4] != b"MThd":
        raise ValueError("not a Standard MIDI File")
    hlen, = struct.unpack(">I", buf[4:8])
    fmt, ntrks, division = struct.unpack(">HHH", buf[8:14])
    if division & 0x8000:
        raise ValueError("SMPTE time division is not supported")
    pattern = Pattern(resolution=division, format=fmt)
    pos = 8 + hlen
    for _ in range(ntrks):
        if buf[pos:pos + 4] != b"MTrk":
            raise ValueError("bad track chunk at byte %d" % pos)
        tlen, = struct.unpack(">I", buf[pos + 4:pos + 8])
        pos += 8
        end = pos + tlen
        track = Track()
        status = None
        while pos < end:
            tick, pos = _read_varlen(buf, pos)
            b = buf[pos]
            if b == 0xFF:
                mtype = buf[pos + 1]
                length, pos = _read_varlen(buf, pos + 2)
                data = buf[pos:pos + length]
                pos += length
                if mtype == 0x51:
                    ev = SetTempoEvent(tick, data)
                elif mtype == 0x2F:
                    ev = EndOfTrackEvent(tick, data)
                else:
                    ev = MetaEvent(tick, data)
                    ev.metacommand = mtype
                track.append(ev)
            elif b in (0xF0, 0xF7):
                length, pos = _read_varlen(buf, pos + 1)
                track.append(SysexEvent(tick, buf[pos:pos + length]))
                pos += length
            else:
                if b & 0x80:
                    status = b
                    pos += 1
                elif status is None:
                    raise ValueError("running status without a status byte")
                n = _CHANNEL_LEN[status >> 4]
                data = buf[pos:pos + n]
                pos += n
                kind = status >> 4
                cls = NoteOnEvent if kind == 0x9 else NoteOffEvent if kind == 0x8 else OtherChannelEvent
                track.append(cls(tick, data, channel=status & 0x0F))
        pos = end
        pattern.append(track)
    return pattern


def read_midifile(fname):
    with open(fname, "rb") as f:
        return parse_bytes(f.read())
