"""(f3) C++ SMF parser (msx_smf_parse, a host function of libmsx.so) against the reference reader's note-event walk on the
reference's own 110 .mid inputs (tests/golden/midi_fixtures.npz, made by tests/golden/make_golden.py), against the
independent pure-Python parser, and on malformed input."""
import os

import numpy as np
import pytest

from musicstyletransfer_b200 import featurise
from musicstyletransfer_b200.MIDIUtil import smf
from musicstyletransfer_b200.MIDIUtil import Melody as M


def test_parser_matches_reference_reader_on_all_fixture_files(golden_dir):
    g = np.load(os.path.join(golden_dir, "midi_fixtures.npz"))
    names = list(g["names"])
    assert len(names) == 110
    n_events = 0
    for name in names:
        info, tracks, ntok = featurise.parse_smf(g["bytes:" + name].tobytes())
        assert info["resolution"] == int(g["res:" + name]) and info["n_tracks"] == int(g["ntracks:" + name]), name
        assert abs(info["bpm"] - float(g["bpm:" + name])) < 1e-9 * max(1.0, info["bpm"]), name
        for ti, (dt, pi, ve) in enumerate(tracks):
            assert np.array_equal(dt, g["dtick:%d:%s" % (ti, name)]), name
            assert np.array_equal(pi, g["pitch:%d:%s" % (ti, name)]) and np.array_equal(ve, g["vel:%d:%s" % (ti, name)]), name
            # tokens-per-track = what A1 makes of the track (midi_io.py:81-89)
            d = dt.astype(np.int64)
            assert int(ntok[ti]) == int(((d + 999) // 1000 * (d > 0)).sum() + len(d)), name
            n_events += len(dt)
    assert n_events > 50000


def test_parser_agrees_with_python_parser_on_synthetic_files(tmp_path):
    rng = np.random.RandomState(3)
    for case in range(20):
        pat = smf.Pattern(resolution=int(rng.choice([96, 120, 220, 480])))
        for _ in range(int(rng.randint(1, 4))):
            tr = smf.Track()
            if rng.rand() < 0.7:
                ev = smf.SetTempoEvent(tick=int(rng.randint(0, 50)))
                ev.set_bpm(float(rng.randint(60, 200)))
                tr.append(ev)
            for _ in range(int(rng.randint(0, 300))):
                tick = int(rng.choice([0, 1, 15, 30, 127, 128, 999, 1000, 2500, 20000]))
                r = rng.rand()
                if r < 0.45:
                    tr.append(smf.NoteOnEvent(tick=tick, pitch=int(rng.randint(0, 128)), velocity=int(rng.randint(0, 128))))
                elif r < 0.9:
                    tr.append(smf.NoteOffEvent(tick=tick, pitch=int(rng.randint(0, 128)), velocity=int(rng.randint(0, 128))))
                else:
                    tr.append(smf.MetaEvent(tick=tick, data=bytes(rng.randint(0, 255, size=int(rng.randint(0, 5))).tolist())))
            tr.append(smf.EndOfTrackEvent(tick=1))
            pat.append(tr)
        path = str(tmp_path / ("c%d.mid" % case))
        smf.write_midifile(path, pat)
        from musicstyletransfer_b200.MIDIUtil.midi_io import note_event_soa
        back = smf.read_midifile(path)
        info, tracks, _ = featurise.parse_smf_file(path)
        assert info["resolution"] == back.resolution and info["n_tracks"] == len(back)
        for tr, (dt, pi, ve) in zip(back, tracks):
            a, b, c = note_event_soa(tr)
            assert np.array_equal(a, dt) and np.array_equal(b, pi) and np.array_equal(c, ve)


def test_writer_roundtrip_through_native_parser(tmp_path):
    from musicstyletransfer_b200.MIDIUtil.midi_io import MelodyWriter
    mel = M.get_melody_from_ids([63, 260, 191, 70, 275, 198])
    mel.resolution, mel.bpm = 120, 90.0
    path = str(tmp_path / "x.mid")
    MelodyWriter().write_to_file(path, mel)
    info, tracks, ntok = featurise.parse_smf_file(path)
    assert info["resolution"] == 120 and abs(info["bpm"] - 90.0) < 1e-3
    dt, pi, ve = tracks[0]
    assert dt.tolist() == [0, 30, 0, 480] and pi.tolist() == [60, 60, 67, 67] and ve.tolist() == [127, 0, 127, 0]
    assert int(ntok[0]) == 6


HDR = b"MThd\x00\x00\x00\x06\x00\x01\x00\x01\x00\x78"


@pytest.mark.parametrize("data,what", [
    (b"RIFFxxxxxxxxxxxxxxxx", "not a Standard MIDI File"),
    (b"MThd\x00\x00", "too short"),
    (b"MThd\x00\x00\x00\x06\x00\x01\x00\x01\xe7\x28MTrk\x00\x00\x00\x00", "SMPTE"),
    (HDR + b"MTrx\x00\x00\x00\x04\x00\x90\x3c\x7f", "bad track chunk"),
    (HDR + b"MTrk\x00\x00\x00\x40\x00\x90\x3c\x7f", "past the end"),
    (HDR + b"MTrk\x00\x00\x00\x03\x00\x3c\x7f", "running status"),
    (HDR + b"MTrk\x00\x00\x00\x04\x00\x90\xc8\x7f", "7-bit"),                 # pitch byte 200
    (HDR + b"MTrk\x00\x00\x00\x03\x00\x90\x3c", "truncated"),
    (HDR + b"MTrk\x00\x00\x00\x05\x00\xff\x51\x7f\x00", "truncated meta"),
])
def test_malformed_files_are_rejected(data, what):
    with pytest.raises(ValueError) as e:
        featurise.parse_smf(data)
    assert what in str(e.value)
    with pytest.raises(ValueError):          # the Python parser rejects the pitch-200 file as well
        if what == "7-bit":
            smf.parse_bytes(data)
        else:
            raise ValueError
