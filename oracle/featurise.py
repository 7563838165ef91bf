"""Oracle: MIDI note events -> event tokens (A1), token rows (A2), piano roll.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain Python / NumPy.
Reference paths are relative to /root/reference/music_style_transfer.
"""
import numpy as np

# --- vocabulary: MIDIUtil/defaults.py:38-58 -------------------------------------------------
MAX_TICKS = 1000                     # defaults.py:38
MIN_TICKS = 0                        # defaults.py:39
NUM_TICKS_IN_A_BIN = 30              # defaults.py:40
NUM_BINS = int((MAX_TICKS - MIN_TICKS) / NUM_TICKS_IN_A_BIN) + 1   # defaults.py:41 -> 34
PAD_ID, SOS_ID, EOS_ID = 0, 1, 2     # defaults.py:44-46
FEATURE_OFFSET = 3                   # defaults.py:48
NOTE_ON_FIRST = FEATURE_OFFSET       # defaults.py:51  (3 .. 130)
NOTE_OFF_FIRST = NOTE_ON_FIRST + 128  # defaults.py:53  (131 .. 258)
TIMESHIFT_FIRST = NOTE_OFF_FIRST + 128  # defaults.py:55 (259 .. 292)
NUM_EVENTS = TIMESHIFT_FIRST + NUM_BINS  # defaults.py:58 -> 293
N_PITCH = 128


def timeshift_id(ticks):
    """create_timeshift_event, MIDIUtil/Melody.py:117-126 (range assert included)."""
    assert MIN_TICKS <= ticks < MAX_TICKS
    binned = int((ticks - MIN_TICKS) / NUM_TICKS_IN_A_BIN)
    assert TIMESHIFT_FIRST + binned <= TIMESHIFT_FIRST + NUM_BINS - 1
    return TIMESHIFT_FIRST + binned


def tokenize_note_events(dtick, pitch, vel):
    """EventBasedMIDIReader._parse_track, MIDIUtil/midi_io.py:70-93, on a stream that holds
    only the note events: ``dtick[i]`` = ticks since the previous *note* event (the reference adds
    every event's tick to ``cur_t`` (:75) but only resets ``prev_t`` on note events (:91), so ticks
    of interleaved non-note events fold into the next note event's delta).
    velocity>0 -> NOTE_ON, velocity==0 -> NOTE_OFF regardless of MIDI event type (:85-89)."""
    ids = []
    for d, p, v in zip(dtick, pitch, vel):
        p, v = int(p) & 0x7F, int(v) & 0x7F                 # MIDI data bytes are 7-bit (malformed input is masked)
        delta = int(d)
        while delta > 0:                                    # :81-83 (lossy modulo quirk)
            ids.append(timeshift_id(delta % MAX_TICKS))
            delta -= MAX_TICKS
        if v > 0:
            ids.append(NOTE_ON_FIRST + int(p))              # Melody.py:109-110
        elif v == 0:
            ids.append(NOTE_OFF_FIRST + int(p))             # Melody.py:113-114
    return ids


def note_events_of_track(track):
    """Fold a parsed SMF track (oracle.smf objects) into the note-event SoA (dtick, pitch, vel)
    exactly as midi_io.py:73-91 walks it."""
    from . import smf
    dt, pi, ve = [], [], []
    prev_t = cur_t = 0
    for ev in track:
        cur_t += ev.tick
        if isinstance(ev, (smf.NoteOnEvent, smf.NoteOffEvent)):
            dt.append(cur_t - prev_t)
            pi.append(ev.data[0])
            ve.append(ev.data[1])
            prev_t = cur_t
    return (np.asarray(dt, dtype=np.int32), np.asarray(pi, dtype=np.uint8),
            np.asarray(ve, dtype=np.uint8))


def read_file_tokens(fname):
    """EventBasedMIDIReader.read_file, midi_io.py:35-68: one token list per surviving track
    (tracks with <10 tokens dropped :60-63, at least one must survive :67).  Returns
    (list_of_id_lists, resolution, list_of_note_event_SoA)."""
    from . import smf
    pattern = smf.read_midifile(fname)
    out, soa = [], []
    for track in pattern:
        ev = note_events_of_track(track)
        ids = tokenize_note_events(*ev)
        if len(ids) < 10:
            continue
        out.append(ids)
        soa.append(ev)
    assert len(out) > 0
    return out, pattern.resolution, soa


# --- A2: MelodyDataset._get_token_arrays, VarAutoEncoder/data.py:133-173 ---------------------
def chunk_rows(melodies_by_class, max_seq_len):
    """melodies_by_class: list (sorted class order) of lists of id-lists.
    Returns float32 tokens [N,L+1], labels [N,L+1], classes [N] (data.py:133-169), including
    the flush-after-every-melody (:149-150), the duplicate row per class (:152-155) and the
    ``labels[:, seq_lens] = EOS`` advanced-indexing quirk (:166-168; NumPy semantics)."""
    L = max_seq_len
    all_tokens, all_classes = [], []
    tokens = None
    for class_idx, melodies in enumerate(melodies_by_class):
        for ids in melodies:
            tokens = np.full((L,), PAD_ID)
            for j, tid in enumerate(ids):
                rel = j % L
                tokens[rel] = tid
                if rel == L - 1:
                    all_tokens.append(tokens)
                    all_classes.append(class_idx)
                    tokens = np.full((L,), PAD_ID)
            all_tokens.append(tokens)
            all_classes.append(class_idx)
        if tokens[0] != PAD_ID:
            all_tokens.append(tokens)
            all_classes.append(class_idx)
    n = len(all_tokens)
    assert n > 0
    data = np.stack(all_tokens).astype(np.float32)
    tok = np.concatenate([np.full((n, 1), SOS_ID, np.float32), data], axis=1)
    seq_lens = (data != PAD_ID).sum(axis=1)                       # data.py:175-179
    labels = np.concatenate([data, np.full((n, 1), PAD_ID, np.float32)], axis=1)
    labels[:, seq_lens] = EOS_ID                                  # data.py:168
    return tok, labels, np.asarray(all_classes, dtype=np.float32)


def preprocess_batch(tokens, labels):
    """MelodyDataset._preprocess_batch, data.py:187-198: seq_lens = #non-PAD incl. SOS; trim
    tokens/labels to max(seq_lens) columns."""
    seq_lens = (tokens != PAD_ID).sum(axis=1).astype(np.float32)
    m = int(seq_lens.max())
    return tokens[:, :m], seq_lens, labels[:, :m]


# --- piano roll (derived spec, SURVEY.md §8(c) "Piano-roll spec"; parity unpinned) -----------
def played_delta(dtick):
    """Ticks the reference's own writer would replay for one note event's time-shift tokens:
    each TIMESHIFT token advances the clock by 30*bin (Melody.py:82-83, midi_io.py:119-127);
    the reader emits ceil(delta/1000) tokens, all with bin (delta%1000)//30 (midi_io.py:81-83)."""
    d = int(dtick)
    if d <= 0:
        return 0
    n_shift = (d + MAX_TICKS - 1) // MAX_TICKS
    return n_shift * NUM_TICKS_IN_A_BIN * ((d % MAX_TICKS) // NUM_TICKS_IN_A_BIN)


def rasterize_sequence(dtick, pitch, vel, resolution, slices_per_quarter, n_slices,
                       max_windows, velocity_roll=False):
    """Token rows + piano roll of ONE note-event sequence.

    Front half = A1 (tokenize_note_events).  Back half: "play" the token stream with the clock
    semantics of MelodyWriter._write_track (midi_io.py:119-127).  Slice index of clock t is
    floor(t*spq/res) (slice width = res/spq ticks: Melody.py:11-16, config.py:37, comment
    midi_io.py:40-49).  NOTE_ON p at slice a: p sounds from a (a re-trigger of a sounding pitch
    closes the old note at a-1 and restarts; velocity roll takes the new velocity).  NOTE_OFF p at
    clock t: the note occupies slices a .. max(a, ceil(t*spq/res)-1); OFF of a silent pitch is
    ignored.  Notes still sounding at the end run to the last slice of the last window.
    Windows: W = min(max_windows, max(1, number of windows touched by the last event));
    events whose slice lies beyond window W-1 are dropped from the roll (tokens keep them).
    Returns (ids, roll uint8 [W, n_slices, 128])."""
    ids = tokenize_note_events(dtick, pitch, vel)
    S = n_slices
    clock = 0
    ev_slice = []
    for d in dtick:
        clock += played_delta(d)
        ev_slice.append((clock * slices_per_quarter) // resolution)
    last_slice = ev_slice[-1] if len(ev_slice) else 0
    W = min(max_windows, last_slice // S + 1)
    total = W * S
    roll = np.zeros((total, N_PITCH), dtype=np.uint8)
    on_slice = [-1] * N_PITCH
    on_val = [0] * N_PITCH
    clock = 0
    for d, p, v, s in zip(dtick, pitch, vel, ev_slice):
        clock += played_delta(d)
        p, v = int(p) & 0x7F, int(v) & 0x7F
        if s >= total:
            break                      # events are time-ordered: everything after is dropped too
        if v > 0:
            if on_slice[p] >= 0:       # re-trigger: old note fills up to s-1 (at least its onset)
                end = max(on_slice[p], s - 1)
                roll[on_slice[p]:end + 1, p] = on_val[p]
            on_slice[p] = s
            on_val[p] = int(v) if velocity_roll else 1
        else:
            if on_slice[p] >= 0:
                end_excl = -((-clock * slices_per_quarter) // resolution)   # ceil
                end = max(on_slice[p], end_excl - 1)
                roll[on_slice[p]:min(end, total - 1) + 1, p] = on_val[p]
                on_slice[p] = -1
    for p in range(N_PITCH):
        if on_slice[p] >= 0:
            roll[on_slice[p]:total, p] = on_val[p]
    return ids, roll.reshape(W, S, N_PITCH)


def rasterize_batch(dtick, pitch, vel, seq_offsets, resolution=120, slices_per_quarter=4,
                    n_slices=64, max_seq_len=64, velocity_roll=False):
    """BASELINE config 2 layout: N independent sequences, one token row [L+1] (SOS + first L
    tokens, PAD-filled) and one roll window [S,128] each.  Also returns the untruncated token
    count per sequence."""
    N = len(seq_offsets) - 1
    tokens = np.full((N, max_seq_len + 1), PAD_ID, dtype=np.int32)
    tokens[:, 0] = SOS_ID
    roll = np.zeros((N, n_slices, N_PITCH), dtype=np.uint8)
    counts = np.zeros((N,), dtype=np.int32)
    for n in range(N):
        a, b = int(seq_offsets[n]), int(seq_offsets[n + 1])
        ids, r = rasterize_sequence(dtick[a:b], pitch[a:b], vel[a:b], resolution,
                                    slices_per_quarter, n_slices, 1, velocity_roll)
        counts[n] = len(ids)
        k = min(len(ids), max_seq_len)
        tokens[n, 1:1 + k] = ids[:k]
        roll[n] = r[0]
    return tokens, roll, counts


def synth_note_events(n_seq=32768, ev_per_seq=32, seed=0):
    """BASELINE config 2 generator: on/off pairs, dtick = 30*U{0..8} with 1 %
    long gaps U{1000..9000}, pitch U{0..127}, on-velocity U{1..127}, off-velocity 0."""
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
