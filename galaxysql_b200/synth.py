"""Synthetic tables of the BASELINE.json shapes: counter-based RNG = splitmix64(seed + row_id),
seed 42, generated on the device for the GPU legs and with numpy for the CPU baseline sample.  Plumbing only
(torch / numpy elementwise ops) — none of this is on the measured path."""
from __future__ import annotations

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

SEED = 42
_STREAM = 0x1000000000


# ------------------------------------------------------------------------------------------------ numpy
def splitmix64_np(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rand_u64_np(n: int, stream: int, seed: int = SEED, start: int = 0) -> np.ndarray:
    return splitmix64_np(np.arange(start, start + n, dtype=np.uint64) + np.uint64(seed) + np.uint64(stream * _STREAM))


def c2_tables_np(build_rows: int, probe_rows: int, key_space: int | None = None, seed: int = SEED):
    """Config 2 shape on the host: (build cols, probe cols).  build.key = a permutation of [0, build_rows)."""
    key_space = key_space or build_rows
    perm = np.argsort(rand_u64_np(build_rows, 0, seed), kind="stable").astype(np.int64)
    build = [perm, (rand_u64_np(build_rows, 1, seed) >> np.uint64(33)).astype(np.int32),
             (rand_u64_np(build_rows, 2, seed) >> np.uint64(33)).astype(np.int32)]
    probe = [(rand_u64_np(probe_rows, 3, seed) % np.uint64(key_space)).astype(np.int64),
             (rand_u64_np(probe_rows, 4, seed) >> np.uint64(33)).astype(np.int32),
             (rand_u64_np(probe_rows, 5, seed) >> np.uint64(33)).astype(np.int32)]
    return build, probe


# ------------------------------------------------------------------------------------------------ torch (device)
def _lsr(x, k):  # logical shift right on int64
    return (x >> k) & ((1 << (64 - k)) - 1)


def splitmix64_t(x):
    """x: int64 tensor (bit pattern of the uint64 counter).  int64 multiply wraps, matching uint64 arithmetic."""
    z = x + (-7046029254386353131)            # 0x9E3779B97F4A7C15
    z = (z ^ _lsr(z, 30)) * (-4658895280553007687)   # 0xBF58476D1CE4E5B9
    z = (z ^ _lsr(z, 27)) * (-7723592293110705685)   # 0x94D049BB133111EB
    return z ^ _lsr(z, 31)


def rand_i64_t(n: int, stream: int, device, seed: int = SEED, start: int = 0, chunk: int = 1 << 27, post=None, out=None):
    """post(bits int64 tensor) -> tensor; generated in chunks to bound temporaries."""
    res = out
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        ctr = torch.arange(start + lo, start + lo + m, dtype=torch.int64, device=device) + (seed + stream * _STREAM)
        bits = splitmix64_t(ctr)
        val = post(bits) if post is not None else bits
        if res is None:
            res = torch.empty(n, dtype=val.dtype, device=device)
        res[lo:lo + m] = val
        del ctr, bits, val
    if res is None:
        res = torch.empty(0, dtype=torch.int64, device=device)
    return res


def _u64_mod(bits, m: int):
    """(uint64 bits) % m for m < 2^31, on int64 tensors."""
    hi = _lsr(bits, 32)
    lo = bits & 0xFFFFFFFF
    return ((hi % m) * ((1 << 32) % m) + lo % m) % m


def _top31(bits):
    return _lsr(bits, 33).to(torch.int32)


def c2_tables_t(build_rows: int, probe_rows: int, device, key_space: int | None = None, key_offset: int = 0,
                seed: int = SEED, probe_start: int = 0, build_perm=None):
    """Config 2 shape on the device.  Returns (build cols, probe cols) as lists of tensors.
    key_space/key_offset let a rank hold one shard of a larger global key space (multi-GPU weak scaling)."""
    key_space = key_space or build_rows
    if build_perm is None:
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        build_perm = torch.randperm(build_rows, generator=g, device=device, dtype=torch.int64)
    build = [build_perm + key_offset if key_offset else build_perm,
             rand_i64_t(build_rows, 1, device, seed, post=_top31),
             rand_i64_t(build_rows, 2, device, seed, post=_top31)]
    probe = [rand_i64_t(probe_rows, 3, device, seed, start=probe_start, post=lambda b: _u64_mod(b, key_space)),
             rand_i64_t(probe_rows, 4, device, seed, start=probe_start, post=_top31),
             rand_i64_t(probe_rows, 5, device, seed, start=probe_start, post=_top31)]
    return build, probe
