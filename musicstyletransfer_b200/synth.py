"""Synthetic workloads named by BASELINE.json / SURVEY.md §8(d) (host-side NumPy generators)."""
import numpy as np

PAD_ID, SOS_ID, EOS_ID = 0, 1, 2          # MIDIUtil/defaults.py:44-46
NOTE_ON_FIRST, NOTE_OFF_FIRST, TIMESHIFT_FIRST = 3, 131, 259   # MIDIUtil/defaults.py:51-56


def token_rows_4_4(n_rows, max_seq_len=64, seed=0):
    """Config 1/3 rows: 4 bars of 16th-grid notes, per slice NOTE_ON(p), TIMESHIFT(bin 1..4), NOTE_OFF(p),
    truncated to L tokens; class 0 = bass pitches 28..62 (defaults.py:34-35), class 1 = guitar 40..88 (:28-29).
    tokens/labels follow MelodyDataset._get_token_arrays (data.py:160-168): tokens = [SOS | data],
    labels = [data | PAD] with EOS written at the column of every occurring row length (all rows full -> column L).
    Returns int32 tokens [N,L+1], seq_lens [N], classes [N], labels [N,L+1]."""
    rng = np.random.RandomState(seed)
    L = max_seq_len
    classes = rng.randint(0, 2, size=n_rows).astype(np.int32)
    n_trip = (L + 2) // 3
    lo = np.where(classes == 0, 28, 40)[:, None]
    hi = np.where(classes == 0, 62, 88)[:, None]
    pitch = (lo + (rng.rand(n_rows, n_trip) * (hi - lo + 1)).astype(np.int64)).astype(np.int64)
    bins = rng.randint(1, 5, size=(n_rows, n_trip))
    data = np.stack([NOTE_ON_FIRST + pitch, TIMESHIFT_FIRST + bins, NOTE_OFF_FIRST + pitch], axis=2)
    data = data.reshape(n_rows, -1)[:, :L].astype(np.int32)
    tokens = np.concatenate([np.full((n_rows, 1), SOS_ID, np.int32), data], axis=1)
    labels = np.concatenate([data, np.full((n_rows, 1), PAD_ID, np.int32)], axis=1)
    labels[:, L] = EOS_ID
    seq_lens = np.full((n_rows,), L + 1, dtype=np.int32)
    return tokens, seq_lens, classes, labels


def note_events(n_seq=32768, ev_per_seq=32, seed=0):
    """Config 2 (1 M note events): on/off pairs, dtick = 30*U{0..8} with 1 % gaps U{1000..9000},
    pitch U{0..127}, on-velocity U{1..127}, off-velocity 0; SoA + int32 offsets."""
    rng = np.random.RandomState(seed)
    E = n_seq * ev_per_seq
    dtick = (30 * rng.randint(0, 9, size=E)).astype(np.int32)
    gaps = rng.rand(E) < 0.01
    dtick[gaps] = rng.randint(1000, 9001, size=int(gaps.sum())).astype(np.int32)
    pitch = np.empty(E, dtype=np.uint8)
    vel = np.empty(E, dtype=np.uint8)
    on_p = rng.randint(0, 128, size=E // 2).astype(np.uint8)
    on_v = rng.randint(1, 128, size=E // 2).astype(np.uint8)
    pitch[0::2] = on_p
    pitch[1::2] = on_p
    vel[0::2] = on_v
    vel[1::2] = 0
    seq_offsets = (np.arange(n_seq + 1, dtype=np.int64) * ev_per_seq).astype(np.int32)
    return dtick, pitch, vel, seq_offsets


def midi_files(n_files=512, ev_per_file=2048, seed=3, resolution=120):
    """Synthetic single-track Standard MIDI Files (bytes) built from note_events(): the input of the end-to-end
    featurisation leg (.mid bytes -> C++ parser -> K1 -> A2 rows -> train steps).  Returns ([bytes], [class index])."""
    import struct
    dtick, pitch, vel, offs = note_events(n_seq=n_files, ev_per_seq=ev_per_file, seed=seed)

    def varlen(v):
        out = [v & 0x7F]
        v >>= 7
        while v:
            out.append((v & 0x7F) | 0x80)
            v >>= 7
        return bytes(reversed(out))
    blobs, classes = [], []
    for i in range(n_files):
        body = bytearray(b"\x00\xff\x51\x03\x07\xa1\x20")                      # tempo 120 bpm
        for e in range(int(offs[i]), int(offs[i + 1])):
            body += varlen(int(dtick[e])) + bytes([0x90 if vel[e] else 0x80, int(pitch[e]) & 0x7F, int(vel[e]) & 0x7F])
        body += b"\x01\xff\x2f\x00"
        blobs.append(b"MThd" + struct.pack(">IHHH", 6, 1, 1, resolution) + b"MTrk" + struct.pack(">I", len(body)) + bytes(body))
        classes.append(i % 2)
    return blobs, classes
