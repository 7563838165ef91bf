"""Standard-MIDI-File reader / writer (host file I/O).  Stands in for the un-pinned third-party ``midi``
(python-midi) module the reference imports (MIDIUtil/midi_io.py:4): ``read_midifile``, ``write_midifile``,
``Pattern`` / ``Track`` and the event classes the reference touches."""
import struct


class Event:
    def __init__(self, tick=0, data=None, channel=0, **kw):
        self.tick = tick
        self.data = list(data) if data is not None else []
        self.channel = channel

    def __repr__(self):
        return "%s(tick=%d, data=%r)" % (type(self).__name__, self.tick, self.data)


class NoteEvent(Event):
    def __init__(self, tick=0, pitch=None, velocity=0, data=None, channel=0):
        super().__init__(tick, data if data is not None else [pitch, velocity], channel)

    @property
    def pitch(self):
        return self.data[0]

    @property
    def velocity(self):
        return self.data[1]


class NoteOnEvent(NoteEvent):
    status = 0x90


class NoteOffEvent(NoteEvent):
    status = 0x80


class OtherChannelEvent(Event):
    status = 0xB0


class MetaEvent(Event):
    metacommand = 0x00


class SetTempoEvent(MetaEvent):
    metacommand = 0x51

    def set_bpm(self, bpm):
        mpqn = int(float(6e7) / bpm)
        self.data = [(mpqn >> 16) & 0xFF, (mpqn >> 8) & 0xFF, mpqn & 0xFF]

    def get_bpm(self):
        return float(6e7) / ((self.data[0] << 16) | (self.data[1] << 8) | self.data[2])


class EndOfTrackEvent(MetaEvent):
    metacommand = 0x2F


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


def _write_varlen(value):
    out = [value & 0x7F]
    value >>= 7
    while value:
        out.append((value & 0x7F) | 0x80)
        value >>= 7
    return bytes(reversed(out))


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
                cls = SetTempoEvent if mtype == 0x51 else EndOfTrackEvent if mtype == 0x2F else MetaEvent
                ev = cls(tick, data)
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
                data = list(buf[pos:pos + n])
                if len(data) != n or any(d & 0x80 for d in data):
                    raise ValueError("malformed channel event at byte %d: data bytes must be 7-bit" % pos)
                pos += n
                kind = status >> 4
                if kind == 0x9:
                    ev = NoteOnEvent(tick, data=data, channel=status & 0x0F)
                elif kind == 0x8:
                    ev = NoteOffEvent(tick, data=data, channel=status & 0x0F)
                else:
                    ev = OtherChannelEvent(tick, data, channel=status & 0x0F)
                    ev.status = status & 0xF0
                track.append(ev)
        pos = end
        pattern.append(track)
    return pattern


def read_midifile(fname):
    with open(fname, "rb") as f:
        return parse_bytes(f.read())


def write_midifile(fname, pattern):
    chunks = []
    for track in pattern:
        body = bytearray()
        for ev in track:
            body += _write_varlen(int(ev.tick))
            if isinstance(ev, MetaEvent):
                body += bytes([0xFF, ev.metacommand]) + _write_varlen(len(ev.data)) + bytes(ev.data)
            elif isinstance(ev, SysexEvent):
                body += bytes([0xF0]) + _write_varlen(len(ev.data)) + bytes(ev.data)
            else:
                body += bytes([ev.status | (ev.channel & 0x0F)]) + bytes(int(d) & 0x7F for d in ev.data)
        chunks.append(b"MTrk" + struct.pack(">I", len(body)) + bytes(body))
    with open(fname, "wb") as f:
        f.write(b"MThd" + struct.pack(">IHHH", 6, pattern.format, len(pattern), pattern.resolution))
        for c in chunks:
            f.write(c)
