"""Synthetic QP workloads for tests and bench.

G1 = batched port of the reference's test generator ``generate_test_QP(n, m, ms, nActive, kappa)``
(reference: interfaces/daqp-julia/test/utils.jl:3-53; MATLAB twin interfaces/daqp-matlab/utils/generate_test_QP.m):
an LDP with a chosen optimal active set and multipliers >= 0 is built first and transformed back to a QP, so the
optimum ``xref`` and the optimal active set are known by construction.

G0 = the probe distribution of BASELINE.md §2 (nearly fully active optima).

Both are pure numpy (counter-based Philox stream: seed = 0x5EED0000 + config id) so the same arrays feed the
oracle, the reference and the CUDA path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEED_BASE = 0x5EED0000


@dataclass
class QPBatch:
    """Homogeneous batch in the packed layout of ``daqp_b200_solve_packed``: every array is C-contiguous,
    leading dimension = problem index. ``A`` holds only the general rows (m - ms)."""

    n: int
    m: int
    ms: int
    H: np.ndarray       # [N, n, n]
    f: np.ndarray       # [N, n]
    A: np.ndarray       # [N, m-ms, n]
    bupper: np.ndarray  # [N, m]
    blower: np.ndarray  # [N, m]
    sense: np.ndarray   # [N, m] int32
    xref: np.ndarray | None = None      # [N, n] known optimum (G1 only)
    active_ref: np.ndarray | None = None  # [N, m] int8: +1 active at upper, -1 at lower, 0 inactive (G1 only)

    @property
    def N(self) -> int:
        return self.H.shape[0]

    def astype(self, dtype) -> "QPBatch":
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dtype)
        return QPBatch(self.n, self.m, self.ms, c(self.H), c(self.f), c(self.A), c(self.bupper), c(self.blower),
                       self.sense.copy(), c(self.xref), self.active_ref)

    def slice(self, lo: int, hi: int) -> "QPBatch":
        s = lambda a: None if a is None else np.ascontiguousarray(a[lo:hi])
        return QPBatch(self.n, self.m, self.ms, s(self.H), s(self.f), s(self.A), s(self.bupper), s(self.blower),
                       s(self.sense), s(self.xref), s(self.active_ref))

    def input_bytes(self) -> int:
        return sum(a.nbytes for a in (self.H, self.f, self.A, self.bupper, self.blower, self.sense))


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=seed))


def generate_g1(N: int, n: int, m: int, ms: int, n_active: int, kappa: float = 100.0, seed: int = SEED_BASE,
                random_nactive: bool = False, near_parallel: tuple | None = None) -> QPBatch:
    """Batched ``generate_test_QP`` (utils.jl:3-53). ``random_nactive`` draws nActive ~ U{0..n_active} per problem
    (config C5: divergent iteration counts). ``near_parallel = (pairs, eps)`` makes `pairs` general rows that are active
    at the constructed optimum copies of another active row plus eps * N(0,1): the optimum sits on nearly dependent
    constraints, which is what drives the reference's ill-conditioning repairs (pivot swaps, refactor-on-exit,
    refinement: daqp.c:28-56, auxiliary.c:379-396)."""
    assert m >= ms and n >= 2 and n_active <= m
    rng = _rng(seed)
    eig = np.ones((N, n))
    eig[:, 1] = kappa
    if n > 2:
        eig[:, 2:] = 1.0 + (kappa - 1.0) * rng.random((N, n - 2))
    Q = np.linalg.qr(rng.standard_normal((N, n, n)))[0]
    sq = np.sqrt(eig)
    T = sq[:, :, None] * np.swapaxes(Q, 1, 2)          # diag(sqrt e) Q'
    Tinv = Q * (1.0 / sq)[:, None, :]                   # Q diag(1/sqrt e)
    H = np.swapaxes(T, 1, 2) @ T

    M = np.empty((N, m, n))
    M[:, :ms, :] = Tinv[:, :ms, :]
    M[:, ms:, :] = rng.standard_normal((N, m - ms, n))

    perm = np.argsort(rng.random((N, m)), axis=1)       # shuffle(1:m)
    pos = np.empty_like(perm)
    np.put_along_axis(pos, perm, np.broadcast_to(np.arange(m), (N, m)), axis=1)
    nact = np.full(N, n_active) if not random_nactive else rng.integers(0, n_active + 1, N)
    nau = (rng.random(N) * (nact + 1)).astype(np.int64)  # rand(0:nActive)
    nau = np.minimum(nau, nact)
    is_up = pos < nau[:, None]
    is_lo = (pos >= nau[:, None]) & (pos < nact[:, None])
    inactive = pos >= nact[:, None]
    sgn = is_up.astype(np.float64) - is_lo.astype(np.float64)

    if near_parallel is not None:
        pairs, eps = near_parallel
        for p in range(N):
            act = np.nonzero((sgn[p] != 0) & (np.arange(m) >= ms))[0]
            for q in range(min(pairs, len(act) // 2)):
                i, j = act[2 * q], act[2 * q + 1]
                M[p, j] = M[p, i] + eps * rng.standard_normal(n)
                sgn[p, j] = sgn[p, i]  # both on the same side, or the pair would pin a slab of width ~eps
        is_up, is_lo = sgn > 0, sgn < 0

    lam = rng.random((N, m)) * (sgn != 0)
    u = -np.einsum("bmn,bm->bn", M, sgn * lam)           # u = -Ma' lam
    Mu = np.einsum("bmn,bn->bm", M, u)
    g1 = 0.01 + rng.random((N, m))
    g2 = 0.01 + rng.random((N, m))
    dupper = np.where(is_up, Mu, np.where(is_lo, Mu + g1, Mu + g1))
    dlower = np.where(is_lo, Mu, np.where(is_up, Mu - g1, Mu - g2))
    del inactive

    v = rng.standard_normal((N, n))
    f = np.einsum("bij,bi->bj", T, v)                    # T' v
    x = np.einsum("bij,bj->bi", Tinv, u - v)             # T \ (u - v)
    A = M[:, ms:, :] @ T
    Mv = np.einsum("bmn,bn->bm", M, v)
    bupper = dupper - Mv
    blower = dlower - Mv
    c = np.ascontiguousarray
    return QPBatch(n, m, ms, c(H), c(f), c(A), c(bupper), c(blower), np.zeros((N, m), dtype=np.int32), c(x),
                   sgn.astype(np.int8))


def generate_miqp(N: int, n: int, m: int, ms: int, nb: int, seed: int = SEED_BASE + 200) -> QPBatch:
    """Batched port of the reference's ``generate_test_MIQP(n, m, ms, nb)`` (interfaces/daqp-julia/test/utils.jl:145-166):
    H = M'M + I, the first nb simple bounds are binary constraints on [0, 1] (sense 16) with a linear term that makes it
    lucrative to leave the origin, and the first general row is the cardinality constraint sum(x[:nb]) <= floor(nb / 2),
    so the relaxation is fractional and branch and bound has to search."""
    assert ms >= nb and m - ms >= 1
    rng = _rng(seed)
    M = rng.standard_normal((N, n, n))
    H = np.swapaxes(M, 1, 2) @ M + np.eye(n)
    A = rng.standard_normal((N, m - ms, n))
    bupper = 20 * rng.random((N, m))
    blower = -20 * rng.random((N, m))
    f = 100 * rng.standard_normal((N, n))
    f[:, :nb] = -np.abs(f[:, :nb])
    bupper[:, :nb] = 1.0
    blower[:, :nb] = 0.0
    sense = np.zeros((N, m), np.int32)
    sense[:, :nb] = 16
    A[:, 0, :] = 0.0
    A[:, 0, :nb] = 1.0
    bupper[:, ms] = np.floor(nb / 2)
    blower[:, ms] = -1e30
    c = np.ascontiguousarray
    return QPBatch(n, m, ms, c(H), c(f), c(A), c(bupper), c(blower), sense)


def generate_g0(N: int, n: int, m: int, seed: int = SEED_BASE + 100) -> QPBatch:
    """Probe distribution of BASELINE.md §2: H = G'G/n + I, f ~ 3 N(0,1), A ~ N(0,1), b = A x0 +- (0.1 + U)."""
    rng = _rng(seed)
    G = rng.standard_normal((N, n, n))
    H = np.swapaxes(G, 1, 2) @ G / n + np.eye(n)
    f = 3.0 * rng.standard_normal((N, n))
    A = rng.standard_normal((N, m, n))
    x0 = 0.1 * rng.standard_normal((N, n))
    Ax0 = np.einsum("bmn,bn->bm", A, x0)
    bupper = Ax0 + 0.1 + rng.random((N, m))
    blower = Ax0 - 0.1 - rng.random((N, m))
    c = np.ascontiguousarray
    return QPBatch(n, m, 0, c(H), c(f), c(A), c(bupper), c(blower), np.zeros((N, m), dtype=np.int32))


# BASELINE.json configs. C4/C5 are widened rows (warm start / fp32 mixed sizes).
CONFIGS = {
    "C1": dict(n=10, m=20, ms=0, n_active=8, N=1),
    "C2": dict(n=20, m=60, ms=0, n_active=16, N=10_000),
    "C3": dict(n=50, m=150, ms=0, n_active=40, N=100_000),
    "C4": dict(n=120, m=400, ms=120, n_active=96, N=50_000),
}


def generate_config(name: str, N: int | None = None, kappa: float = 100.0) -> QPBatch:
    cfg = dict(CONFIGS[name])
    n_default = cfg.pop("N")
    seed = SEED_BASE + int(name[1:])
    return generate_g1(N if N is not None else n_default, kappa=kappa, seed=seed, **cfg)


def generate_g1_torch(N: int, n: int, m: int, ms: int, n_active: int, kappa: float = 100.0, seed: int = SEED_BASE,
                      device="cuda", chunk: int = 8192, random_nactive: bool = False):
    """Same construction as ``generate_g1`` on a torch device (float64), for bench-size batches (100k problems take
    about a second on the GPU instead of a minute in numpy). Returns a dict of contiguous tensors with the packed
    layout of ``daqp_b200_solve_device`` plus ``xref`` / ``active_ref``. Different random stream than the numpy
    generator (same distribution); determinism is per (seed, device type)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    f64 = torch.float64
    mA = m - ms
    out = {"H": torch.empty((N, n, n), dtype=f64, device=dev), "f": torch.empty((N, n), dtype=f64, device=dev),
           "A": torch.empty((N, mA, n), dtype=f64, device=dev), "bupper": torch.empty((N, m), dtype=f64, device=dev),
           "blower": torch.empty((N, m), dtype=f64, device=dev), "xref": torch.empty((N, n), dtype=f64, device=dev),
           "active_ref": torch.empty((N, m), dtype=torch.int8, device=dev)}
    for lo in range(0, N, chunk):
        B = min(chunk, N - lo)
        rnd = lambda *s: torch.rand(*s, dtype=f64, device=dev, generator=g)
        rndn = lambda *s: torch.randn(*s, dtype=f64, device=dev, generator=g)
        eig = torch.ones((B, n), dtype=f64, device=dev)
        eig[:, 1] = kappa
        if n > 2:
            eig[:, 2:] = 1.0 + (kappa - 1.0) * rnd(B, n - 2)
        Q = torch.linalg.qr(rndn(B, n, n))[0]
        sq = eig.sqrt()
        T = sq[:, :, None] * Q.transpose(1, 2)
        Tinv = Q * (1.0 / sq)[:, None, :]
        M = torch.empty((B, m, n), dtype=f64, device=dev)
        M[:, :ms] = Tinv[:, :ms]
        M[:, ms:] = rndn(B, mA, n)
        pos = torch.argsort(torch.argsort(rnd(B, m), dim=1), dim=1)
        nact = torch.full((B,), n_active, device=dev) if not random_nactive else \
            torch.randint(0, n_active + 1, (B,), device=dev, generator=g)
        nau = torch.minimum((rnd(B) * (nact + 1)).long(), nact)
        is_up = pos < nau[:, None]
        is_lo = (pos >= nau[:, None]) & (pos < nact[:, None])
        sgn = is_up.to(f64) - is_lo.to(f64)
        lam = rnd(B, m) * (sgn != 0)
        u = -torch.einsum("bmn,bm->bn", M, sgn * lam)
        Mu = torch.einsum("bmn,bn->bm", M, u)
        g1 = 0.01 + rnd(B, m)
        g2 = 0.01 + rnd(B, m)
        dupper = torch.where(is_up, Mu, Mu + g1)
        dlower = torch.where(is_lo, Mu, torch.where(is_up, Mu - g1, Mu - g2))
        v = rndn(B, n)
        Mv = torch.einsum("bmn,bn->bm", M, v)
        sl = slice(lo, lo + B)
        out["H"][sl] = T.transpose(1, 2) @ T
        out["f"][sl] = torch.einsum("bij,bi->bj", T, v)
        out["A"][sl] = M[:, ms:] @ T
        out["bupper"][sl] = dupper - Mv
        out["blower"][sl] = dlower - Mv
        out["xref"][sl] = torch.einsum("bij,bj->bi", Tinv, u - v)
        out["active_ref"][sl] = sgn.to(torch.int8)
    return out


def torch_to_batch(t: dict, n: int, m: int, ms: int, lo: int = 0, hi: int | None = None) -> QPBatch:
    """Copy a slice of a torch workload to host as a QPBatch (used to hand the SAME problems to the CPU baseline)."""
    hi = t["H"].shape[0] if hi is None else hi
    c = lambda k: np.ascontiguousarray(t[k][lo:hi].cpu().numpy())
    return QPBatch(n, m, ms, c("H"), c("f"), c("A"), c("bupper"), c("blower"), np.zeros((hi - lo, m), np.int32),
                   c("xref"), c("active_ref"))


def soften(b, frac: float, shift: float, seed: int):
    """Mark a random fraction of the constraints SOFT (sense bit 8, reference include/constants.h:84) and move their
    bound pairs by N(0, shift): the hard problem would be infeasible, the soft one ends SOFT_OPTIMAL (exit flag 2)
    with up to n + ns constraints in the working set. Test/bench input only."""
    rng = np.random.default_rng(seed)
    pick = rng.random((b.N, b.m)) < frac
    d = shift * rng.standard_normal((b.N, b.m))
    b.sense[pick] |= 8
    b.bupper[pick] += d[pick]
    b.blower[pick] += d[pick]
    return b


def generate_polyhedra(P: int, n: int, m: int, ms: int = 0, seed: int = 0, redundant_frac: float = 0.4):
    """P random bounded-looking polyhedra {x : [I(ms); A] x <= b} for the minimal-representation path
    (``daqp.minrep``, reference interfaces/daqp-python/daqp.pyx:636-652). Every row is a half-space that contains a ball
    around a random centre c: b_i = a_i'c + |a_i| r_i with r_i in [0.5, 1.5] for the "tight" rows and r_i in [4, 9] for a
    ``redundant_frac`` share of pushed-out rows (most of which end up redundant once enough tight rows exist; whether a
    given row is redundant is for the solver to say). Rows are NOT normalised: |a_i| spans [0.2, 5]. The first ms
    constraints are upper bounds on x_0..x_{ms-1} (unit rows). Returns (A[P, m-ms, n], b[P, m])."""
    rng = np.random.default_rng(seed)
    mA = m - ms
    A = rng.standard_normal((P, mA, n))
    A *= np.exp(rng.uniform(np.log(0.2), np.log(5.0), (P, mA, 1))) / np.linalg.norm(A, axis=2, keepdims=True)
    c = rng.standard_normal((P, n))
    pushed = rng.random((P, m)) < redundant_frac
    r = np.where(pushed, rng.uniform(4.0, 9.0, (P, m)), rng.uniform(0.5, 1.5, (P, m)))
    b = np.empty((P, m))
    b[:, :ms] = c[:, :ms] + r[:, :ms]
    b[:, ms:] = np.einsum("pmn,pn->pm", A, c) + np.linalg.norm(A, axis=2) * r[:, ms:]
    return np.ascontiguousarray(A), np.ascontiguousarray(b)
