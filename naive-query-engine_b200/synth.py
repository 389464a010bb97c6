"""Deterministic synthetic benchmark tables.

mix(x) = splitmix64 finaliser; u(seed, i) = mix(seed + i); unif01 = (u >> 11) * 2^-53.
The same columns can be generated on the host (numpy, vectorised) or directly in
HBM (nqe_synth_column) -- tests check they agree bit for bit.
"""
from __future__ import annotations

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def _chunks(start: int, n: int, step: int = 1 << 24):
    for s in range(0, n, step):
        yield s, min(step, n - s)


def mod_i64(seed: int, start: int, n: int, mod: int, out: np.ndarray = None) -> np.ndarray:
    out = np.empty(n, dtype=np.int64) if out is None else out
    for s, m in _chunks(start, n):
        idx = np.arange(start + s, start + s + m, dtype=np.uint64) + np.uint64(seed)
        out[s:s + m] = (mix64(idx) % np.uint64(mod)).astype(np.int64)
    return out


def unif_f64(seed: int, start: int, n: int, scale: float = 100.0, out: np.ndarray = None) -> np.ndarray:
    out = np.empty(n, dtype=np.float64) if out is None else out
    for s, m in _chunks(start, n):
        idx = np.arange(start + s, start + s + m, dtype=np.uint64) + np.uint64(seed)
        out[s:s + m] = scale * ((mix64(idx) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53)
    return out


def perm_i64(start: int, n: int, mul: int, mod: int, out: np.ndarray = None) -> np.ndarray:
    """(i * mul) % mod -- a bijection on [0, mod) when gcd(mul, mod) == 1."""
    out = np.empty(n, dtype=np.int64) if out is None else out
    for s, m in _chunks(start, n, 1 << 22):
        idx = np.arange(start + s, start + s + m, dtype=np.uint64)
        if (start + n) * mul < 2 ** 63:
            out[s:s + m] = ((idx * np.uint64(mul)) % np.uint64(mod)).astype(np.int64)
        else:
            out[s:s + m] = np.array([(int(i) * mul) % mod for i in idx], dtype=np.int64)
    return out


# column specs of the BASELINE configs: (name, kind, seed, a, b, scale)
#   kind 0: mix(seed+i) % a     kind 1: scale * unif01(mix(seed+i))     kind 2: (i*a) % b
FILTER_TABLE = [("id", 0, 42, 1000, 0, 0.0), ("age", 0, 43, 100, 0, 0.0), ("score", 1, 44, 0, 0, 100.0)]
GROUPBY_TABLE = [("k", 0, 45, 100000, 0, 0.0), ("v", 1, 46, 0, 0, 100.0)]


def join_build_table(n_build: int):
    # k = (i * 7368787) mod n_build (unique keys), a = k mod 100000 is derived by the caller
    return [("k", 2, 0, 7368787, n_build, 0.0)]


def join_probe_table(n_build: int):
    return [("fk", 0, 47, n_build, 0, 0.0), ("b", 1, 48, 0, 0, 100.0)]


def host_column(spec, start: int, n: int, out: np.ndarray = None) -> np.ndarray:
    _, kind, seed, a, b, scale = spec
    if kind == 0:
        return mod_i64(seed, start, n, a, out)
    if kind == 1:
        return unif_f64(seed, start, n, scale, out)
    return perm_i64(start, n, a, b, out)


def device_column(ctx, spec, start: int, n: int, device_ptr: int):
    """Generate the column straight into HBM at `device_ptr` (n 8-byte values)."""
    _, kind, seed, a, b, scale = spec
    ctx.check(ctx.lib.nqe_synth_column(ctx.h, kind, seed, start, n, a, b, scale, device_ptr))
