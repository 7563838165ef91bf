"""Host API of the K1 rasteriser (device featurisation of MIDI note events).

Mirrors what the reference does on the CPU in ``EventBasedMIDIReader._parse_track``
(MIDIUtil/midi_io.py:70-93) and what its writer replays (midi_io.py:119-127), batched on the GPU.
"""
import ctypes

import torch

from . import lib


def rasterize(dtick, pitch, vel, seq_offsets, resolution=120, slices_per_quarter=4, n_slices=64,
              max_seq_len=64, velocity_roll=False, out=None):
    """N independent note-event sequences -> (tokens int32 [N,L+1], roll uint8 [N,S,128], n_tokens int32 [N]).

    dtick int32 [E] (ticks since the previous note event), pitch/vel uint8 [E], seq_offsets int32 [N+1];
    all CUDA tensors.  ``out`` may carry preallocated (tokens, roll, n_tokens)."""
    assert dtick.is_cuda and dtick.dtype == torch.int32 and dtick.is_contiguous()
    assert pitch.dtype == torch.uint8 and vel.dtype == torch.uint8 and seq_offsets.dtype == torch.int32
    assert pitch.is_contiguous() and vel.is_contiguous() and seq_offsets.is_contiguous()
    n = seq_offsets.numel() - 1
    dev = dtick.device
    if out is None:
        tokens = torch.empty((n, max_seq_len + 1), dtype=torch.int32, device=dev)
        roll = torch.empty((n, n_slices, 128), dtype=torch.uint8, device=dev)
        n_tokens = torch.empty((n,), dtype=torch.int32, device=dev)
    else:
        tokens, roll, n_tokens = out
    lib.call("msx_rasterize", lib.ptr(dtick), lib.ptr(pitch), lib.ptr(vel), lib.ptr(seq_offsets),
             ctypes.c_int(n), ctypes.c_int(resolution), ctypes.c_int(slices_per_quarter),
             ctypes.c_int(n_slices), ctypes.c_int(max_seq_len), ctypes.c_int(1 if velocity_roll else 0),
             lib.ptr(tokens), lib.ptr(roll), lib.ptr(n_tokens), lib.stream_ptr())
    return tokens, roll, n_tokens


def rasterize_windows(dtick, pitch, vel, seq_offsets, resolution=120, slices_per_quarter=4, n_slices=64, max_windows=8,
                      velocity_roll=False):
    """Whole tracks as consecutive windows of ``n_slices`` slices (the roll rows of SURVEY.md §8(c): one row per window):
    returns (roll uint8 [N, max_windows, n_slices, 128], n_windows int32 [N]).  Window w of track i is valid for
    w < n_windows[i] = min(max_windows, last event's slice // n_slices + 1); notes sounding at the end of a track run to
    the end of its last window, events beyond window max_windows - 1 are dropped (oracle/featurise.py:rasterize_sequence).
    One K1 launch on a tall tile of max_windows * n_slices slices (<= 1024); the window count comes from the played
    clock of the last event (a segmented sum, plain tensor plumbing)."""
    total = n_slices * max_windows
    assert total <= 1024, "max_windows * n_slices must not exceed 1024 slices (one shared-memory tile per warp)"
    n = seq_offsets.numel() - 1
    _, roll, _ = rasterize(dtick, pitch, vel, seq_offsets, resolution, slices_per_quarter, total, 1, velocity_roll)
    d = dtick.to(torch.int64)
    # played clock of an event (midi_io.py:81-83 replayed by Melody.py:82-83): ceil(d/1000) shifts of 30 * ((d%1000)//30)
    played = torch.where(d > 0, ((d + 999) // 1000) * (30 * ((d % 1000) // 30)), torch.zeros_like(d))
    csum = torch.cat([torch.zeros(1, dtype=torch.int64, device=d.device), torch.cumsum(played, 0)])
    offs = seq_offsets.to(torch.int64)
    clock = csum[offs[1:]] - csum[offs[:-1]]
    last_slice = (clock * slices_per_quarter) // resolution
    n_windows = torch.clamp(last_slice // n_slices + 1, max=max_windows).to(torch.int32)
    return roll.view(n, max_windows, n_slices, 128), n_windows


def tokenize_tracks(soas, device="cuda"):
    """A1 for whole tracks: list of (dtick, pitch, vel) NumPy SoAs -> list of int32 id arrays, untruncated.
    One K1 launch for all tracks (the roll output is a 1-slice dummy window)."""
    import numpy as np
    if not soas:
        return []
    lens = [len(s[0]) for s in soas]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    if offs[-1] == 0:
        return [np.zeros(0, np.int32) for _ in soas]
    cat = lambda i, dt: np.concatenate([np.asarray(s[i], dtype=dt) for s in soas])
    dtick, pitch, vel = cat(0, np.int32), cat(1, np.uint8), cat(2, np.uint8)
    # upper bound of the token count of a track: one note token + ceil(d/1000) shift tokens per event
    bound = max(1, max(int(l + ((np.asarray(s[0], dtype=np.int64) + 999) // 1000).sum()) for l, s in zip(lens, soas)))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    tokens, _, counts = rasterize(t(dtick), t(pitch), t(vel), t(offs), n_slices=1, max_seq_len=bound)
    tokens, counts = tokens.cpu().numpy(), counts.cpu().numpy()
    return [tokens[i, 1:1 + counts[i]].copy() for i in range(len(soas))]


# ------------------------------------------------------------------------------------------------ (f3) SMF bytes -> SoA
class _SmfInfo(ctypes.Structure):
    _fields_ = [("resolution", ctypes.c_int), ("format", ctypes.c_int), ("n_tracks", ctypes.c_int), ("bpm", ctypes.c_double),
                ("n_events", ctypes.c_longlong)]


def parse_smf(data: bytes):
    """Standard MIDI File bytes -> (info dict, [(dtick int32, pitch uint8, vel uint8)] per track, tokens-per-track) through
    the C++ parser in libmsx.so (msx_smf_parse): what midi.read_midifile + the event walk of
    EventBasedMIDIReader._parse_track (MIDIUtil/midi_io.py:35-93) deliver, as the note-event SoA K1 consumes.
    Raises ValueError on malformed files."""
    import numpy as np
    buf = np.frombuffer(data, dtype=np.uint8)
    cap = max(1, buf.size // 3 + 1)
    n_tracks_hdr = int.from_bytes(data[10:12], "big") if len(data) >= 14 else 0
    tcap = max(1, n_tracks_hdr)
    dtick = np.empty(cap, np.int32)
    pitch = np.empty(cap, np.uint8)
    vel = np.empty(cap, np.uint8)
    offs = np.zeros(tcap + 1, np.int32)
    ntok = np.zeros(tcap, np.int32)
    info = _SmfInfo()
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L = lib.load()
    rc = L.msx_smf_parse(vp(buf), ctypes.c_longlong(buf.size), ctypes.c_longlong(cap), ctypes.c_int(tcap), vp(dtick),
                         vp(pitch), vp(vel), vp(offs), vp(ntok), ctypes.byref(info))
    if rc != 0:
        raise ValueError(L.msx_last_error().decode())
    tracks = [(dtick[offs[t]:offs[t + 1]].copy(), pitch[offs[t]:offs[t + 1]].copy(), vel[offs[t]:offs[t + 1]].copy())
              for t in range(info.n_tracks)]
    return ({"resolution": info.resolution, "format": info.format, "n_tracks": info.n_tracks, "bpm": info.bpm,
             "n_events": int(info.n_events)}, tracks, ntok[:info.n_tracks].copy())


def parse_smf_file(path):
    with open(path, "rb") as f:
        return parse_smf(f.read())


# ------------------------------------------------------------------------------------------------ A2 on the device
def build_rows(tokens, n_tokens, track_class, class_start, max_seq_len, col0=1):
    """K1's per-track token streams -> the dataset rows of MelodyDataset._get_token_arrays (VarAutoEncoder/data.py:133-173)
    on the device.  tokens int32 [N, ld] with the ids of track t at tokens[t, col0 : col0 + n_tokens[t]] (K1's untruncated
    output: col0 = 1 skips its SOS column), n_tokens int32 [N], track_class int32 [N] (tracks grouped by class in the
    reference's iteration order), class_start int32 [C + 1]; all CUDA tensors.
    Returns (tokens int32 [R, L+1], labels int32 [R, L+1], classes int32 [R], seq_lens int32 [R]); one 4-byte device->host
    read (the row count) sizes the outputs — a one-off per dataset."""
    N = n_tokens.numel()
    C = class_start.numel() - 1
    L = int(max_seq_len)
    dev = tokens.device
    assert tokens.dtype == torch.int32 and tokens.is_contiguous() and n_tokens.dtype == torch.int32
    assert track_class.dtype == torch.int32 and class_start.dtype == torch.int32
    row_start = torch.empty(N + 1, dtype=torch.int32, device=dev)
    dup_row = torch.empty(C, dtype=torch.int32, device=dev)
    dup_src = torch.empty(C, dtype=torch.int32, device=dev)
    len_present = torch.empty(L + 1, dtype=torch.int32, device=dev)
    i32 = ctypes.c_int
    lib.call("msx_rows_plan", lib.ptr(n_tokens), lib.ptr(class_start), i32(N), i32(C), i32(L), lib.ptr(row_start),
             lib.ptr(dup_row), lib.ptr(dup_src), lib.ptr(len_present), lib.stream_ptr())
    R = int(row_start[N].item())
    out_tok = torch.empty((R, L + 1), dtype=torch.int32, device=dev)
    out_lab = torch.empty((R, L + 1), dtype=torch.int32, device=dev)
    out_cls = torch.empty((R,), dtype=torch.int32, device=dev)
    out_len = torch.empty((R,), dtype=torch.int32, device=dev)
    lib.call("msx_rows_build", lib.ptr(tokens), ctypes.c_longlong(tokens.stride(0)), i32(col0), lib.ptr(n_tokens),
             lib.ptr(track_class), i32(N), i32(C), i32(L), lib.ptr(row_start), lib.ptr(dup_row), lib.ptr(dup_src),
             lib.ptr(len_present), lib.ptr(out_tok), lib.ptr(out_lab), lib.ptr(out_cls), lib.ptr(out_len), lib.stream_ptr())
    return out_tok, out_lab, out_cls, out_len


def gather_batch(rows, index, t_out):
    """rows = build_rows(...) output; index int32 [B] (CUDA) -> (tokens [B, t_out], labels [B, t_out], classes [B],
    seq_lens [B]) int32: the batch _preprocess_batch hands to the step (data.py:187-198), trimmed to t_out columns."""
    tok, lab, cls, lens = rows
    B = index.numel()
    dev = tok.device
    bt = torch.empty((B, t_out), dtype=torch.int32, device=dev)
    bl = torch.empty((B, t_out), dtype=torch.int32, device=dev)
    bc = torch.empty((B,), dtype=torch.int32, device=dev)
    bn = torch.empty((B,), dtype=torch.int32, device=dev)
    i32 = ctypes.c_int
    lib.call("msx_rows_gather_batch", lib.ptr(tok), lib.ptr(lab), lib.ptr(cls), lib.ptr(lens), lib.ptr(index), i32(B),
             i32(tok.shape[1]), i32(int(t_out)), lib.ptr(bt), lib.ptr(bl), lib.ptr(bc), lib.ptr(bn), lib.stream_ptr())
    return bt, bl, bc, bn


def tokenize_tracks_device(soas, device="cuda"):
    """A1 for whole tracks, results left on the device: (tokens int32 [N, bound + 1] with SOS in column 0, n_tokens int32 [N])."""
    import numpy as np
    lens = [len(s[0]) for s in soas]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cat = lambda i, dt: (np.concatenate([np.asarray(s[i], dtype=dt) for s in soas]) if offs[-1] else np.zeros(0, dt))
    dtick, pitch, vel = cat(0, np.int32), cat(1, np.uint8), cat(2, np.uint8)
    bound = max(1, max([int(l + ((np.asarray(s[0], dtype=np.int64) + 999) // 1000).sum()) for l, s in zip(lens, soas)] or [1]))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    if offs[-1] == 0:
        n = len(soas)
        tok = torch.zeros((n, bound + 1), dtype=torch.int32, device=device)
        tok[:, 0] = 1
        return tok, torch.zeros((n,), dtype=torch.int32, device=device)
    tokens, _, counts = rasterize(t(dtick), t(pitch), t(vel), t(offs), n_slices=1, max_seq_len=bound)
    return tokens, counts
