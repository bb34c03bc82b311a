"""Host replica of the device dropout stream (csrc/train_simt.cuh hash_uniform): lets tests and
data-parallel runs reproduce the SELU-dropout keep mask of any (seed, global element index).
The reference's dropout is unseeded (selu.py:55), so any stream is as faithful as another."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash_uniform(seed, idx):
    """uniform [0,1) float32 for uint64 indices `idx` (array) under `seed`"""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.asarray(idx, np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def keep_mask(seed, n_sites, rate, width=336, site0=0):
    """(n_sites, width) float32 mask = floor(keep + U) exactly as k_dropout_fwd draws it (selu.py:54-56)"""
    idx = np.arange(site0 * width, (site0 + n_sites) * width, dtype=np.uint64)
    u = hash_uniform(seed, idx)
    return np.floor(np.float32(1.0 - rate) + u).reshape(n_sites, width).astype(np.float32)
