"""CPU emulation (numpy, fp32) of the sorted walk over relu crossings that the fixed-grid kernels use to keep the
piecewise-linear head coefficients current (csrc/slode_mlp_kernels.cuh, PlEval::build / advance_half): keys =
early-biased crossing coordinates with the unit index in the low mantissa bits, sorted once; an evaluation examines
the due candidates in order and confirms each with the exact fp32 gate test.  The property the reverse sweep's
bookkeeping relies on: after every evaluation the walked gate pattern equals the dense test
``fma(w1t_j, t, c_j) >= 0`` for every unit, on increasing and decreasing sweeps, uniform and ragged grids."""
import numpy as np
import pytest

f32 = np.float32


def walk_patterns(w1t, c, evals, dirsign, index_bits):
    H = w1t.shape[0]
    imask = np.uint32((1 << index_bits) - 1)
    never = np.uint32(0x7F800000)
    with np.errstate(divide="ignore"):
        rinv = np.where(w1t == 0, f32(0), f32(-1) / w1t).astype(f32)
    t_start = f32(evals[0])
    dirsign = f32(dirsign)
    cur = ~np.signbit((w1t * t_start + c).astype(f32))
    ts = (c * rinv).astype(f32)
    slack = (f32(4e-7) * (np.abs(ts) + abs(t_start))).astype(f32)
    u = (dirsign * (ts - t_start)).astype(f32)
    ub = np.maximum((u * f32(0.99998474) - slack).astype(f32), f32(0))
    pending = (rinv != 0) & (u > -(f32(64) * slack + f32(1e-30))) & (ub < f32(3e38))
    keys = np.where(pending, (ub.view(np.uint32) & ~imask) | np.arange(H, dtype=np.uint32), never)
    ks = np.append(np.sort(keys), never)
    pos = 0
    out = [cur.copy()]
    for te in evals[1:]:
        te = f32(te)
        uq = f32(dirsign * (te - t_start))
        while uq >= ks[pos:pos + 1].view(f32)[0]:
            j = int(ks[pos] & imask)
            post = bool(dirsign * w1t[j] > 0)
            now = not np.signbit(f32(w1t[j] * te + c[j]))
            if cur[j] != post:
                if now != post:
                    break
                cur[j] = post
            pos += 1
        out.append(cur.copy())
    return out


def eval_times(grid, method):
    ev = []
    for t0, t1 in zip(grid[:-1], grid[1:]):
        dt = f32(t1 - t0)
        if method == "euler":
            ev += [t0]
        elif method == "midpoint":
            ev += [t0, f32(t0 + f32(0.5) * dt)]
        else:
            ev += [t0, f32(t0 + dt * f32(1 / 3)), f32(t0 + dt * f32(2 / 3))]
    return ev + ([grid[-1]] if method == "rk4" else [])


@pytest.mark.parametrize("H,index_bits", [(25, 5), (32, 5)])
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_walk_reproduces_the_dense_gate_patterns(H, index_bits, method):
    rng = np.random.default_rng(7)
    for case in range(40):
        w1t = rng.uniform(-0.25, 0.25, H).astype(f32)
        if case % 5 == 0:
            w1t[rng.integers(H)] = 0.0                      # a unit that never flips
        c = (rng.normal(size=H) * rng.choice([0.05, 0.6, 3.0])).astype(f32)
        if case % 2:
            grid = np.cumsum(rng.uniform(0.05, 0.9, 60)).astype(f32) - f32(rng.choice([0.0, 20.0, 1000.0]))
        else:
            grid = np.arange(0, 40, dtype=f32)
        fwd = eval_times(grid, method)
        for evals, dirsign in ((fwd, 1.0), (fwd[::-1], -1.0)):    # forward kernel, reverse sweep
            if grid[-1] < grid[0]:
                dirsign = -dirsign
            walked = walk_patterns(w1t, c, evals, dirsign, index_bits)
            for te, pat in zip(evals, walked):
                exact = ~np.signbit((w1t * f32(te) + c).astype(f32))
                assert (pat == exact).all(), (case, method, dirsign, float(te), np.nonzero(pat != exact)[0])
