"""ctypes wrapper of oracle/raster.c (TEST INFRASTRUCTURE ONLY; built by __graft_entry__.build())."""
import ctypes
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_raster.so")


def available():
    return os.path.exists(_SO)


def rasterize_batch(dtick, pitch, vel, seq_offsets, resolution=120, slices_per_quarter=4, n_slices=64, max_seq_len=64,
                    velocity_roll=False, threads=1):
    lib = ctypes.CDLL(_SO)
    n = len(seq_offsets) - 1
    dtick = np.ascontiguousarray(dtick, np.int32)
    pitch = np.ascontiguousarray(pitch, np.uint8)
    vel = np.ascontiguousarray(vel, np.uint8)
    offs = np.ascontiguousarray(seq_offsets, np.int32)
    tokens = np.empty((n, max_seq_len + 1), np.int32)
    roll = np.empty((n, n_slices, 128), np.uint8)
    counts = np.empty((n,), np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.oracle_rasterize(p(dtick), p(pitch), p(vel), p(offs), n, resolution, slices_per_quarter, n_slices, max_seq_len,
                         1 if velocity_roll else 0, p(tokens), p(roll), p(counts), int(threads))
    return tokens, roll, counts
