"""numpy restatement of the framework's counter-based generator (Philox4x32-10).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference has no
counter-based generator (it uses the global legacy MT19937, app.py:702), so
this file restates the *new* generator's published algorithm -- Salmon, Moraes,
Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Philox4x32
with 10 rounds -- and the framework's own counter layout and uniform ->
exponential / normal transforms, so that in-kernel-RNG runs can be checked
value-by-value and not only statistically.  Pinned by the Random123
known-answer vectors in ``tests/test_oracle.py``.  **Parity with the reference
is unpinned by construction** (different generator; distributional equivalence
to ``np.random.dirichlet(ones)`` is what the statistical tests check).

Counter layout (128-bit counter, 64-bit key):
    c0, c1 = low / high 32 bits of the GLOBAL unit index (portfolio or path)
    c2     = sub-counter: rejection attempt (portfolios) or time step (paths)
    c3     = (stream << 24) | block;  stream 1 = weights, 2 = path normals
    key    = low / high 32 bits of the seed

Uniforms per unit (asset i of a portfolio, output i of a path step):
    float64  one 32-bit word each: word i%4 of block i//4, U = 1 - x * 2^-32
    float32  24-bit fields: field i = bits [24 i, 24 i + 23) of the concatenated blocks 0, 1, 2, ...
             (word 0 of block 0 lowest), i.e. four 23-bit fractions per three words -- 16 uniforms cost
             3 Philox calls instead of 4 (`philox_fields` in mcp_device.cuh); U = 1 - field * 2^-23
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_WEIGHTS = 1
STREAM_NORMALS = 2


def philox4x32(c0, c1, c2, c3, k0, k1, rounds: int = 10):
    """Vectorised Philox4x32-R.  Inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _raw_outputs(index, sub, n_outputs, stream, seed, rounds=10):
    """uint32 array (len(index), n_outputs): output j = word j%4 of block j//4."""
    index = np.asarray(index, dtype=np.uint64)
    lo = (index & MASK32)[:, None]
    hi = (index >> np.uint64(32))[:, None]
    nblk = (n_outputs + 3) // 4
    blocks = np.arange(nblk, dtype=np.uint64)[None, :]
    c3 = (np.uint64(stream) << np.uint64(24)) | blocks
    sub = np.asarray(sub, dtype=np.uint64)
    sub = sub[:, None] if sub.ndim == 1 else sub
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32(lo, hi, sub, c3, seed & 0xFFFFFFFF, seed >> 32, rounds)
    out = np.stack(r, axis=-1).reshape(index.shape[0], nblk * 4)
    return out[:, :n_outputs]


def _fields24(index, sub, n_fields, stream, seed, rounds=10):
    """uint32 array (len(index), n_fields) of 23-bit fractions: the float32 kernels' uniform fields."""
    n_trip = (n_fields + 3) // 4
    w = _raw_outputs(index, sub, 3 * n_trip + ((-3 * n_trip) % 4), stream, seed, rounds).astype(np.uint64)
    a, b, c = w[:, 0:3 * n_trip:3], w[:, 1:3 * n_trip:3], w[:, 2:3 * n_trip:3]
    f = np.empty((w.shape[0], 4 * n_trip), dtype=np.uint64)
    f[:, 0::4] = a
    f[:, 1::4] = (a >> np.uint64(24)) | (b << np.uint64(8))
    f[:, 2::4] = (b >> np.uint64(16)) | (c << np.uint64(16))
    f[:, 3::4] = c >> np.uint64(8)
    return (f[:, :n_fields] & np.uint64(0x7FFFFF)).astype(np.uint32)


def exponentials(index, attempt, n_assets, seed, dtype="float32", rounds=10):
    """Base-2 exponentials e = -log2(U), U in (0, 1].

    float32: U = 1 - field * 2^-23 (23 mantissa bits, built on the GPU with one LOP3 as a
    float in [1,2) and one subtraction; 24-bit fields, see the module docstring);
    float64: U = 1 - x * 2^-32.  The ln 2 factor cancels in the normalisation w = e / sum(e).
    """
    sub = np.broadcast_to(np.asarray(attempt, dtype=np.uint64), np.shape(index))
    if dtype == "float32":
        u = 1.0 - _fields24(index, sub, n_assets, STREAM_WEIGHTS, seed, rounds).astype(np.float64) * 2.0 ** -23
    else:
        x = _raw_outputs(index, sub, n_assets, STREAM_WEIGHTS, seed, rounds)
        u = 1.0 - x.astype(np.float64) * 2.0 ** -32
    return -np.log2(u)


def dirichlet_weights(first_index, n_portfolios, n_assets, seed, dtype="float32",
                      min_weights=None, max_weights=None, max_tries=100, keep_last=False, rounds=10):
    """Flat-Dirichlet weights for global indices [first, first+P) with bounds rejection.

    Mirrors the reference's loop (app.py:699-707): attempt t = 0..max_tries-1, accept the
    first draw inside [min, max]; returns (W, valid, tries_used).
    """
    idx = np.arange(first_index, first_index + n_portfolios, dtype=np.uint64)
    W = np.zeros((n_portfolios, n_assets))
    valid = np.zeros(n_portfolios, dtype=bool)
    pending = np.ones(n_portfolios, dtype=bool)
    lo = None if min_weights is None else np.asarray(min_weights, dtype=np.float64)
    hi = None if max_weights is None else np.asarray(max_weights, dtype=np.float64)
    for t in range(max_tries):
        sel = np.nonzero(pending)[0]
        if sel.size == 0:
            break
        e = exponentials(idx[sel], t, n_assets, seed, dtype, rounds)
        w = e / e.sum(axis=1, keepdims=True)
        if dtype == "float32":
            # bounds are compared in the kernel's arithmetic type
            wc = w.astype(np.float32)
            ok = np.ones(sel.size, dtype=bool)
            if lo is not None:
                ok &= np.all(wc >= lo.astype(np.float32), axis=1)
            if hi is not None:
                ok &= np.all(wc <= hi.astype(np.float32), axis=1)
        else:
            ok = np.ones(sel.size, dtype=bool)
            if lo is not None:
                ok &= np.all(w >= lo, axis=1)
            if hi is not None:
                ok &= np.all(w <= hi, axis=1)
        W[sel] = w
        valid[sel[ok]] = True
        pending[sel[ok]] = False
    if keep_last:
        valid[:] = True
    return W, valid


def normals(first_index, n_paths, n_steps, n_assets, seed, dtype="float32", rounds=10):
    """Standard normals Z[m, s, i] by Box-Muller on output pairs (2k, 2k+1).

    U1 in (0,1] from output 2k, angle fraction f in [0,1) from output 2k+1:
    z[2k] = r cos(theta), z[2k+1] = r sin(theta), r = sqrt(-2 ln U1), theta = pi (2f - 1)
    (the angle is centred on 0 so the MUFU sin/cos approximations stay in [-pi, pi)).
    float32 uses the 23-bit fractions of the 24-bit fields, float64 32-bit words (same convention as `exponentials`).
    """
    idx = np.arange(first_index, first_index + n_paths, dtype=np.uint64)
    n_even = (n_assets + 1) // 2 * 2
    Z = np.empty((n_paths, n_steps, n_assets))
    for s in range(n_steps):
        if dtype == "float32":
            f = _fields24(idx, np.full(n_paths, s, dtype=np.uint64), n_even, STREAM_NORMALS, seed, rounds).astype(np.float64) * 2.0 ** -23
        else:
            x = _raw_outputs(idx, np.full(n_paths, s, dtype=np.uint64), n_even, STREAM_NORMALS, seed, rounds)
            f = x.astype(np.float64) * 2.0 ** -32
        u1 = 1.0 - f[:, 0::2]
        th = np.pi * (2.0 * f[:, 1::2] - 1.0)
        r = np.sqrt(-2.0 * np.log(u1))
        z = np.empty((n_paths, n_even))
        z[:, 0::2] = r * np.cos(th)
        z[:, 1::2] = r * np.sin(th)
        Z[:, s, :] = z[:, :n_assets]
    return Z
