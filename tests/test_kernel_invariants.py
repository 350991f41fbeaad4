"""CPU checks of the algorithmic rewrites the CUDA kernels rely on (hypothesis; no GPU, no oracle).  Each test restates, in a
few lines of Python, the reference's sequential rule and the data-parallel form a kernel uses instead, and asserts that they
agree on random and adversarial inputs:

* k_project      : runs of survivors with non-decreasing canonical x are pushed in bulk, out-of-order points replay the scalar
                   pop / skip / push of velo.h:351-368
* k_assoc_search : the binary search of velo.h:404-412 without early `continue`s (both neighbours read, lo / hi by selects); the
                   x table of a ring (gaps by the lanes, head / tail by the whole warp) + the prefix count over the keypoint's
                   own bucket == that search on a ring whose x never decreases
* k_icp_pass     : the running best as one unsigned 64-bit (bits of d2, index) minimum == smallest d2, ties to the lower index;
                   the branch-free insertion of per-ring results into the two smallest keys == the reference's streaming top-2 over
                   ascending rings (velo.h:825-848), in ANY visiting order, and a bound taken from two seed points never cuts it off;
                   the range bucket taken from the float's bits is monotone; the partition of the queries into blocks / runs
                   covers every query exactly once for any number of CTAs
"""
import numpy as np
from hypothesis import given, settings, strategies as st

SET = dict(max_examples=200, deadline=None)
f32 = np.float32


# ------------------------------------------------------------------------------------------------ occlusion stack
def stack_reference(cx, z):
    """velo.h:351-368 on the in-FOV survivors of one ring, in ring order; returns the indices left on the stack"""
    out = []
    for i in range(len(cx)):
        while out and cx[i] < cx[out[-1]] and z[i] < z[out[-1]]:
            out.pop()
        if out and cx[i] < cx[out[-1]] and z[i] > z[out[-1]]:
            continue
        out.append(i)
    return out


def stack_ordered_runs(cx, z, chunk=32):
    """k_project: per chunk of survivors, the longest ordered prefix is pushed at once, the first violator takes the scalar path"""
    out = []
    for c0 in range(0, len(cx), chunk):
        rem = list(range(c0, min(len(cx), c0 + chunk)))
        while rem:
            run = []
            for k, i in enumerate(rem):
                pred = rem[k - 1] if k else (out[-1] if out else None)
                if pred is not None and cx[i] < cx[pred]:
                    break
                run.append(i)
            out.extend(run)                       # no pop, no skip possible: every point of the run is pushed
            rem = rem[len(run):]
            if rem:
                i = rem.pop(0)                    # the first out-of-order survivor: reference's pop / skip / push
                while out and cx[i] < cx[out[-1]] and z[i] < z[out[-1]]:
                    out.pop()
                if out and cx[i] < cx[out[-1]] and z[i] > z[out[-1]]:
                    continue
                out.append(i)
    return out


@settings(**SET)
@given(seed=st.integers(0, 10**6), n=st.integers(0, 200), disorder=st.sampled_from([0.0, 0.05, 0.5, 1.0]), quant=st.booleans())
def test_ordered_run_push_equals_sequential_stack(seed, n, disorder, quant):
    rng = np.random.default_rng(seed)
    cx = np.sort(rng.uniform(-1, 1, n))
    swap = rng.random(n) < disorder
    cx[swap] = rng.uniform(-1, 1, swap.sum())                      # out-of-order points (depth discontinuities)
    z = rng.uniform(2, 40, n)
    if quant:                                                      # equal x and equal z (neither pop nor skip fires on equality)
        cx = np.round(cx * 8) / 8; z = np.round(z / 4) * 4
    cx, z = cx.astype(f32), z.astype(f32)
    for chunk in (32, 5, 1):
        assert stack_ordered_runs(cx, z, chunk) == stack_reference(cx, z)


# ------------------------------------------------------------------------------------------------ bracket search
def search_reference(px, kx):
    lo, hi = 0, len(px) - 2
    while lo <= hi:
        mid = (lo + hi) >> 1
        if px[mid] > kx:
            hi = mid - 1; continue
        if px[mid + 1] <= kx:
            lo = mid + 1; continue
        return mid
    return -1


def search_uniform(px, kx):
    lo, hi, mid, found = 0, len(px) - 2, 0, False
    while lo <= hi and not found:
        mid = (lo + hi) >> 1
        a, b = px[mid], px[mid + 1]
        left = a > kx; right = (not left) and (b <= kx)
        hi = mid - 1 if left else hi
        lo = mid + 1 if right else lo
        found = (not left) and (not right)
    return mid if found else -1


@settings(**SET)
@given(seed=st.integers(0, 10**6), n=st.integers(2, 300), sorted_=st.booleans(), quant=st.booleans())
def test_uniform_binary_search_equals_reference(seed, n, sorted_, quant):
    rng = np.random.default_rng(seed)
    px = rng.uniform(-1, 1, n)
    if sorted_: px = np.sort(px)                                   # the stack output is not strictly sorted: test both
    if quant: px = np.round(px * 16) / 16                          # duplicated x (hazard H4)
    px = px.astype(f32)
    for kx in list(rng.uniform(-1.1, 1.1, 20).astype(f32)) + [px[0], px[-1], px[n // 2]]:
        assert search_uniform(px, kx) == search_reference(px, kx)


# ------------------------------------------------------------------------------------------------ x table of a ring
NB = 128


def xbucket(x, xmin, xscale):
    return int(min(max(int((f32(x) - f32(xmin)) * f32(xscale)), 0), NB - 1))


def lut_build(px, xmin, xscale):
    """k_assoc_search: lane i > 0 fills the buckets between its predecessor's and its own, the warp fills the head and the tail"""
    cnt = len(px)
    lut = [None] * (NB + 1)
    b = [xbucket(x, xmin, xscale) for x in px]
    for i in range(1, cnt):
        for k in range(b[i - 1] + 1, b[i] + 1): lut[k] = i
    bfirst, blast = (b[0], b[-1]) if cnt else (NB, NB)
    for k in range(0, bfirst + 1): lut[k] = 0
    for k in range(blast + 1, NB + 1): lut[k] = cnt
    return lut


def lut_search(px, lut, kx, xmin, xscale):
    """two table reads + the prefix count over the bucket, four points at a time"""
    xb = xbucket(kx, xmin, xscale)
    j, j1 = lut[xb], lut[xb + 1]
    while True:
        j0 = j
        for t in range(4):
            jj = max(min(j0 + t, j1 - 1), 0)
            j += int(j0 + t < j1 and px[jj] <= kx)
        if j != j0 + 4: break
    mid = j - 1
    return mid if 0 <= mid <= len(px) - 2 else None


@settings(**SET)
@given(seed=st.integers(0, 10**6), n=st.integers(2, 300), span=st.sampled_from([0.05, 0.4, 1.0]), dup=st.booleans())
def test_x_table_search_equals_reference_on_a_monotone_ring(seed, n, span, dup):
    rng = np.random.default_rng(seed)
    xmin, xmax = f32(-0.9), f32(0.9)
    xscale = f32(NB) / (xmax - xmin)
    lo = rng.uniform(-1.0, 1.0 - span)                                 # rings that cross only part of the image, or leave it
    px = np.sort(rng.uniform(lo, lo + span, n)).astype(f32)
    if dup: px = (np.round(px * 64) / 64).astype(f32)                  # duplicated x (velo.h:360-366 pushes z ties)
    lut = lut_build(px, xmin, xscale)
    for k in range(NB + 1):                                            # the definition: first index whose bucket is >= k
        first = next((i for i in range(n) if xbucket(px[i], xmin, xscale) >= k), n)
        assert lut[k] == first
    for kx in np.concatenate([rng.uniform(-1.1, 1.1, 30), px[rng.integers(0, n, 10)]]).astype(f32):
        want = search_reference(px, kx)
        got = lut_search(px, lut, kx, xmin, xscale)
        if want >= 0:                                                  # px[mid] <= kx < px[mid+1] has one solution on a monotone ring
            assert got == want and px[got] <= kx < px[got + 1]
        else:
            assert got is None


# ------------------------------------------------------------------------------------------------ 64-bit running best
@settings(**SET)
@given(seed=st.integers(0, 10**6), n=st.integers(1, 80), ties=st.booleans())
def test_u64_key_minimum_is_smallest_distance_then_lowest_index(seed, n, ties):
    rng = np.random.default_rng(seed)
    d2 = rng.uniform(0, 0.5, n).astype(f32)
    if ties: d2 = (np.round(d2 * 8) / 8).astype(f32)               # exact distance ties (includes +0.0)
    idx = rng.permutation(n).astype(np.uint32)                     # `sorted` is in azimuth-bin order, not index order
    thr_excl = f32(0.3)
    best = np.uint64(thr_excl.view(np.uint32)) << np.uint64(32)    # scan_init: (threshold bits, index 0)
    for k in range(n):
        key = (np.uint64(d2[k].view(np.uint32)) << np.uint64(32)) | np.uint64(idx[k])
        best = min(best, key)
    ok = d2 < thr_excl
    found = (int(best) >> 32) < int(thr_excl.view(np.uint32))
    assert found == bool(ok.any())
    if found:
        m = d2[ok].min()
        assert np.uint32(int(best) >> 32).view(f32) == m
        assert (int(best) & 0xFFFFFFFF) == idx[ok & (d2 == m)].min()


# ------------------------------------------------------------------------------------------------ top-2 rings
def top2_reference(per_ring):
    """velo.h:825-848: rings in ascending order, strict `<` against the running best two distances"""
    inf = float("inf")
    di, dj, si, sj, ni, nj = inf, inf, -1, -1, -1, -1
    for s, (d, n) in enumerate(per_ring):
        if d is None: continue
        if d < di: dj, sj, nj = di, si, ni; di, si, ni = d, s, n
        elif d < dj: dj, sj, nj = d, s, n
    return (si, ni), (sj, nj)


@settings(**SET)
@given(seed=st.integers(0, 10**6), n_rings=st.integers(1, 70), ties=st.booleans(), seeded=st.booleans())
def test_branch_free_top2_insertion_equals_streaming_top2(seed, n_rings, ties, seeded):
    """Every ring is visited once and hands in its nearest point inside the threshold as the key (bits of d2 | ring | index); the
    kernel keeps the two smallest keys with lo = min(k, ki), hi = max(k, ki), ki = lo, kj = min(kj, hi), in whatever order the
    rings come (nearest elevation first, several mask levels).  With a seed bound (the farther of two real points of different
    rings) rings whose nearest point lies beyond the bound are never visited: the result must not change."""
    rng = np.random.default_rng(seed)
    IDX_BITS = 20
    d = rng.uniform(0, 0.5, n_rings).astype(f32)
    if ties: d = (np.round(d * 8) / 8).astype(f32)
    has = rng.random(n_rings) < 0.7                                # rings with a point inside the threshold
    idx = rng.integers(0, 2000, n_rings)
    per_ring = [(float(d[s]), int(idx[s])) if has[s] else (None, None) for s in range(n_rings)]
    (si, ni), (sj, nj) = top2_reference(per_ring)
    INF = (1 << 64) - 1
    key = lambda s: (int(d[s].view(np.uint32)) << 32) | (s << IDX_BITS) | int(idx[s])
    bound = None
    if seeded and has.sum() >= 2:                                  # any two rings with points: the farther one bounds the runner-up
        a, b = rng.choice(np.nonzero(has)[0], 2, replace=False)
        bound = max(d[a], d[b])
    ki = kj = INF
    for s in rng.permutation(n_rings):
        if not has[s] or (bound is not None and d[s] > bound): continue
        k = key(s); lo, hi = min(k, ki), max(k, ki)
        ki, kj = lo, min(kj, hi)
    ring = lambda k: -1 if k == INF else (k & 0xFFFFFFFF) >> IDX_BITS
    pidx = lambda k: -1 if k == INF else k & ((1 << IDX_BITS) - 1)
    assert (ring(ki), pidx(ki)) == (si, ni)
    assert (ring(kj), pidx(kj)) == (sj, nj)


# ------------------------------------------------------------------------------------------------ window half-width
def test_asin_upper_bound_survives_the_approximate_arithmetic():
    """asin_ub(x) = x / sqrt(1 - x^2) * (1 + 4e-6) + 3e-5 sizes the azimuth window / elevation tolerance of a query (csrc/velo_icp.cu).
    Its input b / D comes from the 2-ulp hardware square root and division and its rsqrt is the 2-ulp MUFU one: with every one of
    those errors at its worst (1e-6 relative on x, 5e-7 on the rsqrt, float rounding of each step) the result must still exceed
    asin(x) by the 1e-5 rad kept for the polynomial atan2_q, for every x the kernel does not already treat as 'full circle'."""
    x = np.concatenate([np.linspace(0.0, 0.999, 200001), np.geomspace(1e-9, 1e-2, 2001)])
    worst = np.inf
    for ex in (-1e-6, 1e-6):
        for er in (-5e-7, 5e-7):
            xa = (x * (1 + ex)).astype(f32)                                            # what the kernel computed for the true x
            t = (f32(1.0) - (xa * xa).astype(f32)).astype(f32)
            rs = ((1.0 / np.sqrt(t.astype(np.float64))) * (1 + er)).astype(f32)
            ub = (((xa * rs).astype(f32) * f32(1.0 + 4e-6)).astype(f32) + f32(3e-5)).astype(f32)
            worst = min(worst, float((ub.astype(np.float64) - np.arcsin(x) - 1e-5).min()))
    assert worst > 0.0, worst


# ------------------------------------------------------------------------------------------------ range buckets
def rg_bucket(rho, bits=6, rg_min=2.0, n_buckets=256):
    r = np.maximum(np.asarray(rho, f32), f32(rg_min))
    b = (r.view(np.int32) - f32(rg_min).view(np.int32)) >> (23 - bits)
    return np.minimum(b, n_buckets - 1)


def test_range_bucket_is_monotone_and_covers_the_sensor_range():
    rho = np.sort(np.concatenate([np.linspace(0, 200, 200001), np.random.default_rng(0).uniform(0, 40, 100000)])).astype(f32)
    b = rg_bucket(rho)
    assert np.all(np.diff(b) >= 0) and b.min() == 0 and b.max() == 255
    assert rg_bucket(f32(1.0)) == 0 and rg_bucket(f32(np.inf)) == 255
    assert len(np.unique(b[(rho >= 2) & (rho < 32)])) == 256 and rg_bucket(f32(31.9)) == 255   # 64 per octave over 2..32 m


# ------------------------------------------------------------------------------------------------ blocks / runs
@settings(**SET)
@given(Q=st.integers(0, 200000), ctas=st.integers(1, 300))
def test_run_partition_covers_every_query_once(Q, ctas):
    WARPS, CHUNKS = 8, 4
    BLOCK = 32 * CHUNKS * WARPS
    nblk = (Q + BLOCK - 1) // BLOCK
    blk_per = (nblk + ctas - 1) // ctas if nblk else 0
    seen = np.zeros(Q, np.int32); records = set()
    for cta in range(ctas):
        blk0, blk1 = cta * blk_per, min(nblk, cta * blk_per + blk_per)
        for run in range(max(0, blk1 - blk0) * WARPS):
            blk, rw = blk0 + run // WARPS, run % WARPS
            records.add(blk * WARPS + rw)
            for ci in range(CHUNKS):
                qb = blk * BLOCK + (rw + ci * WARPS) * 32
                if qb >= Q: break
                seen[qb:min(Q, qb + 32)] += 1
    assert np.all(seen == 1)
    assert records == set(range(nblk * WARPS))                     # exactly the records k_neq_reduce adds
