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
