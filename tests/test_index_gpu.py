"""GPU parity of the landmark-index path: CUDA (through the C-ABI) vs the oracle and the reference's
golden vectors.  Bar: indices bit-exact; scores within 1e-3 relative (scale-aware) -- in
practice they are the same bits because both sides round one fp64-accumulated dot."""
from pathlib import Path

import numpy as np
import pytest

import synth
from oracle import avl_oracle as O

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"
SCORE_RTOL = 1e-3  # the tolerance north_star states for fp32 similarity scores


def assert_scores_close(got, ref, feat, q):
    floor = (np.linalg.norm(feat, axis=1)[:, None] * np.linalg.norm(q, axis=1)[None, :]) / np.sqrt(feat.shape[1])
    assert np.all(np.abs(got - ref) <= SCORE_RTOL * np.maximum(np.abs(ref), floor))


@pytest.fixture(scope="module")
def eng(lib):
    from avlmaps_b200 import engine

    return engine


@pytest.mark.parametrize("name", ["c1_10k_q2", "4k_q64", "3k_d768_q9", "odd_1001_d100_q3"])
def test_golden_reference_vectors(eng, name):
    """BASELINE config 1 and friends: identical mask/argmax to the reference's own output."""
    g = np.load(G / f"index_{name}.npz")
    feat, q = synth.index_inputs(int(g["n"]), int(g["d"]), int(g["nq"]), int(g["seed"]))
    m = eng.DeviceMap(feat)
    am = m.argmax(q)
    assert am.dtype == np.int32 and np.array_equal(am, g["argmax"])       # vlmap.py:123
    assert np.array_equal(am == 0, g["mask0"])                              # vlmap.py:124
    sc = m.scores(q)
    assert_scores_close(sc, g["scores"], feat, q)                           # vs the reference's float32 BLAS
    assert np.array_equal(sc, O.scores(feat, q))                            # vs the oracle: same bits
    assert np.array_equal(np.argmax(sc, axis=1), am)                        # fused argmax == argmax of dense scores
    m.close()


@pytest.mark.parametrize("n,d,nq,k,normalize,use_scale", [
    (1000, 512, 1, 1, False, False),
    (1000, 512, 2, 16, False, False),
    (5000, 512, 9, 16, False, False),
    (70_001, 512, 64, 16, False, False),
    (33_333, 512, 65, 5, True, True),
    (20_000, 512, 256, 16, False, False),   # 256 queries: cta_group::2, B resident across the SM pair
    (50_001, 448, 200, 16, True, True),
    (9_000, 768, 33, 8, True, False),
    (9_000, 1024, 32, 16, True, True),     # AudioCLIP-like: unit rows, scale 100
    (777, 100, 3, 128, False, False),       # D not a multiple of 64, k > typical
    (50, 512, 3, 16, False, False),         # fewer rows than k
])
def test_topk_and_argmax_vs_oracle(eng, n, d, nq, k, normalize, use_scale):
    feat, q = synth.index_inputs(n, d, nq, seed=n % 97, unit_rows=(d == 1024))
    scale = np.random.default_rng(5).uniform(0.5, 100.0, nq).astype(np.float32) if use_scale else None
    ref = O.scores(feat, q, scale=scale, normalize=normalize)
    m = eng.DeviceMap(feat)
    idx, val = m.topk(q, k, scale=scale, normalize_map=normalize)
    ri, rv = O.topk(ref, k)
    assert idx.dtype == np.int64 and np.array_equal(idx, ri)
    assert np.array_equal(val, rv)
    am = m.argmax(q, scale=scale, normalize_map=normalize)
    assert np.array_equal(am, O.argmax(ref))
    assert np.array_equal(m.scores(q, scale=scale, normalize_map=normalize), ref)
    m.close()


@pytest.mark.parametrize("cg", [1, 2])
def test_tensor_core_screen_matches_bf16_model(eng, cg):
    """The raw tcgen05 output equals the exact product of the bf16-rounded operands (fp32 accumulate);
    cg = cta_group of the kernel (2: the queries are split over an SM pair)."""
    feat, q = synth.index_inputs(3000, 512, 48, seed=3)

    def bf16(x):
        b = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
        return ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)

    m = eng.DeviceMap(feat)
    got = m.screen_scores(q, cta_group=cg)
    ref = bf16(feat).astype(np.float64) @ bf16(q).astype(np.float64).T
    scale = np.linalg.norm(feat, axis=1)[:, None] * np.linalg.norm(q, axis=1)[None, :]
    assert np.max(np.abs(got - ref) / scale) < 2e-6
    m.close()


@pytest.mark.parametrize("cg", [1, 2])
def test_f16_operand_mode(eng, cg):
    """AVL_MAP_F16: fp16 tensor-core operands.  (a) the raw screen equals the exact product of the fp16-rounded
    operands; (b) argmax / top-k are the same bits as with bf16 operands and as the oracle, with ~8x fewer rows
    inside the error band; (c) a value beyond the fp16 range makes the map fall back to bf16 by itself."""
    feat, q = synth.index_inputs(30_000, 512, 48, seed=3)
    m16, mb = eng.DeviceMap(feat, operand="f16"), eng.DeviceMap(feat)
    assert m16.operand == "f16" and mb.operand == "bf16"
    got = m16.screen_scores(q, cta_group=cg)
    ref = feat.astype(np.float16).astype(np.float64) @ q.astype(np.float16).astype(np.float64).T
    scale = np.linalg.norm(feat, axis=1)[:, None] * np.linalg.norm(q, axis=1)[None, :]
    assert np.max(np.abs(got - ref) / scale) < 2e-6
    sc = O.scores(feat, q)
    a16 = m16.argmax(q, want_stats=True)
    f16_flagged = m16.last_stats["n_flagged"]
    ab = mb.argmax(q, want_stats=True)
    assert np.array_equal(a16, ab) and np.array_equal(a16, O.argmax(sc))
    assert f16_flagged * 4 < mb.last_stats["n_flagged"]
    i16, v16 = m16.topk(q, 16)
    ri, rv = O.topk(sc, 16)
    assert np.array_equal(i16, ri) and np.array_equal(v16, rv)
    m16.close()
    mb.close()
    big = feat.copy()
    big[7, 3] = 1.0e5                        # > 65504
    mf = eng.DeviceMap(big, operand="f16")
    assert mf.operand == "bf16"
    assert np.array_equal(mf.argmax(q), O.argmax(O.scores(big, q)))
    mf.close()


def test_tiny_map_never_returns_rows_past_the_end(eng):
    """Fewer sample groups than k gives the threshold -inf; rows of the last tile past the end of the map must
    still be rejected (they were only harmless while freshly allocated device memory happened to be zero).
    Dirty the allocator first, then query maps whose row count is far from a multiple of the 128-row tile."""
    import torch

    junk = torch.full((64 << 20,), 3.0e4, device="cuda")   # 256 MB of large values, freed -> recycled by cudaMalloc
    del junk
    torch.cuda.empty_cache()
    for n, nq, k in ((50, 3, 16), (129, 200, 16), (5, 2, 8)):
        feat, q = synth.index_inputs(n, 512, nq, seed=n)
        feat -= 40.0                                        # every real score is far below 0: stale rows would win
        m = eng.DeviceMap(feat)
        sc = O.scores(feat, q)
        idx, val = m.topk(q, k)
        ri, rv = O.topk(sc, k)
        assert idx.max() < n and np.array_equal(idx, ri) and np.array_equal(val, rv)
        assert np.array_equal(m.argmax(q), O.argmax(sc))
        m.close()


def test_ties_resolve_to_lowest_index(eng):
    feat, q = synth.index_inputs(4096, 512, 4, seed=1)
    feat[100:110] = feat[7]          # ten exact copies of row 7
    feat[2000] = feat[7]
    q[1] = q[0]                      # two identical queries: argmax must pick the lower id
    m = eng.DeviceMap(feat)
    ref = O.scores(feat, q)
    idx, val = m.topk(q, 32)
    ri, rv = O.topk(ref, 32)
    assert np.array_equal(idx, ri) and np.array_equal(val, rv)
    am = m.argmax(q)
    assert np.array_equal(am, O.argmax(ref)) and not np.any(am == 1)
    m.close()


def test_massive_ties_take_the_exact_fallback(eng):
    """All rows identical: every row is a candidate, the candidate list overflows, the exact dense
    fallback must still return rows 0..k-1."""
    row = synth.index_inputs(1, 512, 1, seed=2)[0]
    feat = np.repeat(row, 20_000, axis=0)
    q = synth.index_inputs(1, 512, 3, seed=4)[1]
    m = eng.DeviceMap(feat)
    idx, val = m.topk(q, 8)
    assert np.array_equal(idx, np.tile(np.arange(8), (3, 1)))
    assert m.last_stats["n_fallback_queries"] == 3
    assert np.array_equal(val, O.topk(O.scores(feat, q), 8)[1])
    m.close()


@pytest.mark.parametrize("n,d,nq,k,normalize", [(30_000, 100, 21, 128, False), (25_000, 1024, 11, 5, True)])
def test_fallback_scores_groups_of_queries_per_pass(eng, n, d, nq, k, normalize):
    """Near-duplicate rows (a few prototypes + one-ulp noise): every query's top-k is a near-tie far inside the 16-bit band,
    so every query overflows into the device-side exact fallback, which scores up to 8 flagged queries per pass over the
    map (fewer when k or D is large).  Ids, score bits and order must equal the oracle's."""
    rng = np.random.default_rng(6)
    protos = rng.standard_normal((4, d)).astype(np.float32)
    feat = protos[rng.integers(0, 4, n)] * np.float32(3.0)
    feat *= (1.0 + 1e-7 * rng.integers(-3, 4, (n, 1))).astype(np.float32)
    q = synth.index_inputs(1, d, nq, seed=8)[1]
    scale = rng.uniform(0.5, 2.0, nq).astype(np.float32)
    ref = O.topk(O.scores(feat, q, scale=scale, normalize=normalize), k)
    m = eng.DeviceMap(feat)
    idx, val = m.topk(q, k, scale=scale, normalize_map=normalize)
    assert m.last_stats["n_fallback_queries"] == nq
    assert np.array_equal(idx, ref[0]) and np.array_equal(val, ref[1])
    m.close()


def test_zero_rows_and_empty_map(eng):
    feat, q = synth.index_inputs(2000, 512, 5, seed=9)
    feat[::7] = 0.0
    ref = O.scores(feat, q, normalize=True)
    m = eng.DeviceMap(feat)
    idx, val = m.topk(q, 16, normalize_map=True)
    ri, rv = O.topk(ref, 16)
    assert np.array_equal(idx, ri) and np.array_equal(val, rv)
    m.close()
    e = eng.DeviceMap(np.zeros((0, 512), np.float32))
    idx, val = e.topk(q, 4)
    assert np.all(idx == -1) and np.all(np.isneginf(val))
    assert e.argmax(q).shape == (0,)
    e.close()


def test_deterministic(eng):
    feat, q = synth.index_inputs(50_000, 512, 64, seed=12)
    m = eng.DeviceMap(feat)
    a1, a2 = m.argmax(q), m.argmax(q)
    i1, v1 = m.topk(q, 16)
    i2, v2 = m.topk(q, 16)
    assert np.array_equal(a1, a2) and np.array_equal(i1, i2) and np.array_equal(v1, v2)
    m.close()


def test_shard_merge_equals_single_map(eng):
    """Slab sharding property: merging per-shard top-k equals the top-k of the whole map."""
    from avlmaps_b200.sharded import merge_topk

    feat, q = synth.index_inputs(30_000, 512, 16, seed=21)
    whole = eng.DeviceMap(feat)
    wi, wv = whole.topk(q, 16)
    cuts = [0, 7_000, 7_100, 19_999, 30_000]
    parts_i, parts_v = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        s = eng.DeviceMap(feat[a:b])
        i, v = s.topk(q, 16)
        parts_i.append(np.where(i >= 0, i + a, -1))
        parts_v.append(v)
        s.close()
    mi, mv = merge_topk(np.stack(parts_i), np.stack(parts_v), 16)
    assert np.array_equal(mi, wi) and np.array_equal(mv, wv)
    whole.close()


def test_vector_topk_and_fusion(eng):
    rng = np.random.default_rng(3)
    v = rng.standard_normal(1_000_003).astype(np.float32)
    v[rng.integers(0, v.size, v.size // 3)] = 1.0
    gi, gv = eng.topk_vector(v, 16)
    ri, rv = O.topk_vector(v, 16)
    assert np.array_equal(gi, ri) and np.array_equal(gv, rv)
    # BASELINE config 3, reduced: LSeg-512 x AudioCLIP-1024, min-max, product, top-16
    n = 40_000
    fv, qv = synth.index_inputs(n, 512, 6, seed=30)
    fa, qa = synth.index_inputs(n, 1024, 6, seed=31, unit_rows=True)
    sa = np.full(6, 100.0, np.float32)  # clamp ceiling of the AudioCLIP logit scale (audioclip.py:174)
    mv_, ma_ = eng.DeviceMap(fv), eng.DeviceMap(fa)
    for combine in (O.FUSE_PRODUCT, O.FUSE_MAX, O.FUSE_SUM):
        gi, gh = eng.fuse_topk(mv_, qv, ma_, qa, 16, scale_b=sa, combine=combine)
        ri, rh = O.fuse_topk(O.scores(fv, qv), O.scores(fa, qa, scale=sa), combine, 16)
        assert np.array_equal(gi, ri)
        assert np.allclose(gh, rh, rtol=0, atol=2e-7)
    mv_.close()
    ma_.close()


def test_device_pointer_path_matches_host_path(eng):
    import torch

    feat, q = synth.index_inputs(20_000, 512, 40, seed=5)
    m = eng.DeviceMap(torch.from_numpy(feat).cuda())
    qi = torch.from_numpy(q).cuda()
    i_d, v_d = m.topk(qi, 16)
    a_d = m.argmax(qi)
    i_h, v_h = m.topk(q, 16)
    assert np.array_equal(i_d.cpu().numpy(), i_h) and np.array_equal(v_d.cpu().numpy(), v_h)
    assert np.array_equal(a_d.cpu().numpy(), m.argmax(q))
    m.close()


def test_capi_argument_errors(eng, lib):
    import ctypes as C

    from avlmaps_b200 import _lib as L

    feat, q = synth.index_inputs(100, 512, 4, seed=0)
    m = eng.DeviceMap(feat)
    oi, ov = np.empty((4, 200), np.int64), np.empty((4, 200), np.float32)
    assert lib.avl_sim_topk(m._h, L.np_ptr(q), 4, None, 0, 200, L.np_ptr(oi), L.np_ptr(ov), 0, None, None) == 2
    assert lib.avl_sim_topk(m._h, L.np_ptr(q), 0, None, 0, 4, L.np_ptr(oi), L.np_ptr(ov), 0, None, None) == 2
    bad_scale = np.array([1, -1, 1, 1], np.float32)
    assert lib.avl_sim_topk(m._h, L.np_ptr(q), 4, L.np_ptr(bad_scale), 0, 4, L.np_ptr(oi), L.np_ptr(ov), 0, None, None) == 2
    with pytest.raises(ValueError):
        m.topk(q[:, :100], 4)
    m.close()


def test_full_size_c2_properties(eng):
    """BASELINE config 2 at full size (1M x 512, Q = 64) through size-independent properties: the
    argmax equals the oracle on a row sample; every returned top-k score is the exact score of its
    row; no sampled row beats the k-th score; the two halves merge to the whole."""
    import torch

    from avlmaps_b200.sharded import merge_topk

    n, d, nq, k = 1_000_000, 512, 64, 16
    g = torch.Generator(device="cuda").manual_seed(1)
    feat_t = torch.randn((n, d), device="cuda", generator=g) * (torch.rand((n, 1), device="cuda", generator=g) * 13.5 + 0.7)
    q_t = torch.randn((nq, d), device="cuda", generator=g)
    q_t = q_t / q_t.norm(dim=1, keepdim=True)
    m = eng.DeviceMap(feat_t)
    am = m.argmax(q_t, want_stats=True).cpu().numpy()
    idx, val = m.topk(q_t, k)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    q = q_t.cpu().numpy()
    rows = np.unique(np.concatenate([np.random.default_rng(0).integers(0, n, 20_000), idx.reshape(-1)]))
    sub = feat_t[torch.from_numpy(rows).cuda()].cpu().numpy()
    ref = O.scores(sub, q)
    assert np.array_equal(am[rows], O.argmax(ref))
    pos = {r: i for i, r in enumerate(rows)}
    for j in range(nq):
        exact = np.array([ref[pos[r], j] for r in idx[j]])
        assert np.array_equal(val[j], exact)
        assert np.all(np.diff(val[j]) <= 0)
        beaten = ref[:, j] > val[j, -1]
        assert set(rows[beaten]).issubset(set(idx[j]))
    h = n // 2
    parts = []
    for a, b in ((0, h), (h, n)):
        s = eng.DeviceMap(feat_t[a:b])
        i, v = s.topk(q_t, k)
        parts.append((np.where(i.cpu().numpy() >= 0, i.cpu().numpy() + a, -1), v.cpu().numpy()))
        s.close()
    mi, mv = merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), k)
    assert np.array_equal(mi, idx) and np.array_equal(mv, val)
    m.close()


def test_headline_size_properties(eng):
    """The headline shape (4 194 304 x 512, 256 queries, top-16; cta_group::2 kernel) through size-independent
    properties: (a) for 3 of the queries the result equals the exact top-16 over ALL rows (exact dense column,
    a kernel that is bit-identical to the oracle at small sizes); (b) for all 256 queries every returned score
    is the oracle's score of its row, sorted, and no row of a 20k random sample beats the k-th score."""
    import torch

    n, d, nq, k = 4_194_304, 512, 256, 16
    g = torch.Generator(device="cuda").manual_seed(3)
    feat_t = torch.empty((n, d), dtype=torch.float32, device="cuda")
    for r0 in range(0, n, 1 << 20):
        r1 = min(n, r0 + (1 << 20))
        feat_t[r0:r1] = torch.randn((r1 - r0, d), device="cuda", generator=g)
        feat_t[r0:r1] *= torch.rand((r1 - r0, 1), device="cuda", generator=g) * 0.6 + 0.03
    q_t = torch.randn((nq, d), device="cuda", generator=g)
    q_t = (q_t / q_t.norm(dim=1, keepdim=True)).contiguous()
    m = eng.DeviceMap(feat_t)
    idx_t, val_t = m.topk(q_t, k)
    assert m.last_stats["cta_group"] == 2 and m.last_stats["n_fallback_queries"] == 0
    idx, val = idx_t.cpu().numpy(), val_t.cpu().numpy()
    pick = [0, 101, 255]
    cols = m.scores(q_t[pick].contiguous()).cpu().numpy()          # (n, 3) exact canonical scores
    for c, j in enumerate(pick):
        ri, rv = O.topk_vector(np.ascontiguousarray(cols[:, c]), k)
        assert np.array_equal(idx[j], ri) and np.array_equal(val[j], rv)
    q = q_t.cpu().numpy()
    rows = np.unique(np.concatenate([np.random.default_rng(1).integers(0, n, 20_000), idx.reshape(-1)]))
    ref = O.scores(feat_t[torch.from_numpy(rows).cuda()].cpu().numpy(), q)
    pos = {r: i for i, r in enumerate(rows)}
    for j in range(nq):
        assert np.array_equal(val[j], np.array([ref[pos[r], j] for r in idx[j]]))
        assert np.all(np.diff(val[j]) <= 0)
        assert set(rows[ref[:, j] > val[j, -1]]).issubset(set(idx[j]))
    m.close()


def test_heat_from_mask_bit_exact(eng):
    """get_heatmap_from_mask_3d (visualize_utils.py:29-49): golden vector of the reference + a larger oracle case."""
    g = np.load(G / "heat_n600.npz")
    h = eng.heat_from_mask_3d(g["pos"], g["mask"], float(g["cell_size"]), float(g["decay_rate"]))
    assert h.dtype == np.float32 and np.array_equal(h, g["heat"])
    rng = np.random.default_rng(1)
    pos = rng.integers(0, 300, (30_000, 3)).astype(np.int32)
    mask = rng.uniform(size=30_000) < 0.02
    want = O.heatmap_from_mask_3d(pos, mask, 0.05, 0.01)
    assert np.array_equal(eng.heat_from_mask_3d(pos, mask, 0.05, 0.01), want)
    from avlmaps_b200 import _lib as L

    with pytest.raises(L.AvlError, match="selects no voxel"):
        eng.heat_from_mask_3d(pos, np.zeros(30_000, bool))


@pytest.mark.parametrize("decay", [0.1, 0.01, 0.004, 0.0009])
def test_heat_windowed_search_equals_brute_force(eng, monkeypatch, decay):
    """The heat is clipped to 0 beyond cell_size / decay cells, so the kernel only looks for targets inside that
    ball (bitmap + offsets in ascending distance); same bits as the brute-force kernel and as the oracle.
    decay 0.1 (AVLMap.index_object's default): radius 0.5, only the targets are hot; 0.0009: ball larger than the
    target set, the library picks brute force by itself."""
    rng = np.random.default_rng(7)
    n = 200_000
    pos = np.stack([rng.integers(-20, 400, n), rng.integers(0, 400, n), rng.integers(0, 30, n)], 1).astype(np.int32)
    mask = rng.uniform(size=n) < 0.01
    got = eng.heat_from_mask_3d(pos, mask, 0.05, decay)
    monkeypatch.setenv("AVL_HEAT_BRUTE", "1")
    brute = eng.heat_from_mask_3d(pos, mask, 0.05, decay)
    monkeypatch.delenv("AVL_HEAT_BRUTE")
    assert np.array_equal(got, brute)
    monkeypatch.setenv("AVL_HEAT_BITMAP", "1")   # second strategy: target bitmap + per-voxel walk of the ball
    assert np.array_equal(eng.heat_from_mask_3d(pos, mask, 0.05, decay), brute)
    monkeypatch.delenv("AVL_HEAT_BITMAP")
    sub = rng.integers(0, n, 3000)
    want = O.heatmap_from_mask_3d(np.concatenate([pos[sub], pos[mask]]), np.concatenate([np.zeros(3000, bool) | mask[sub], np.ones(int(mask.sum()), bool)]), 0.05, decay)
    assert np.array_equal(got[sub], want[:3000])
    if decay >= 0.004:
        assert (got > 0).sum() > mask.sum() or decay == 0.1


def _fuse_both_paths(eng, *args, **kw):
    import os

    got = eng.fuse_topk(*args, **kw)              # tcgen05 screens + interval propagation + exact re-score
    os.environ["AVL_FUSE_EXACT"] = "1"
    try:
        want = eng.fuse_topk(*args, **kw)         # exact dense columns
    finally:
        del os.environ["AVL_FUSE_EXACT"]
    return got, want


@pytest.mark.parametrize("normalize", [False, True])
def test_fusion_through_the_screen_equals_exact_path(eng, normalize):
    """BASELINE config 3 shape (LSeg-512 + AudioCLIP-1024, 32 + 32 queries, scale 100, top-16), reduced to 300k rows:
    the screened path returns the same ids and the same heat bits as the exact path, and both match the oracle."""
    n, pairs = 300_000, 32
    fv, qv = synth.index_inputs(n, 512, pairs, seed=40)
    fa, qa = synth.index_inputs(n, 1024, pairs, seed=41, unit_rows=True)
    sa = np.full(pairs, 100.0, np.float32)
    mv_, ma_ = eng.DeviceMap(fv), eng.DeviceMap(fa)
    for combine in (O.FUSE_PRODUCT, O.FUSE_MAX, O.FUSE_SUM):
        (gi, gh), (wi, wh) = _fuse_both_paths(eng, mv_, qv, ma_, qa, 16, scale_b=sa, combine=combine,
                                               normalize_a=normalize, normalize_b=False)
        assert np.array_equal(gi, wi) and np.array_equal(gh, wh)
    ri, rh = O.fuse_topk(O.scores(fv, qv, normalize=normalize), O.scores(fa, qa, scale=sa), O.FUSE_PRODUCT, 16)
    (gi, gh), _ = _fuse_both_paths(eng, mv_, qv, ma_, qa, 16, scale_b=sa, combine=O.FUSE_PRODUCT, normalize_a=normalize)
    assert np.array_equal(gi, ri) and np.allclose(gh, rh, rtol=0, atol=2e-7)
    mv_.close()
    ma_.close()


def test_fusion_degenerate_and_tied_inputs_fall_back_to_the_exact_path(eng):
    """A constant column (max == min -> NaN heat, like numpy) and massive exact ties overflow the candidate lists;
    the call is then answered by the exact path, so both paths still agree."""
    n, pairs = 50_000, 4
    fv, qv = synth.index_inputs(n, 64, pairs, seed=50)
    fa, qa = synth.index_inputs(n, 128, pairs, seed=51, unit_rows=True)
    fv[:] = fv[0]                       # every row identical: every column of modality a is constant
    fa[1000:] = fa[1000]                # 49k exact duplicates in modality b
    mv_, ma_ = eng.DeviceMap(fv), eng.DeviceMap(fa)
    (gi, gh), (wi, wh) = _fuse_both_paths(eng, mv_, qv, ma_, qa, 8, combine=O.FUSE_SUM)
    assert np.array_equal(gi, wi) and np.array_equal(gh, wh, equal_nan=True)
    mv_.close()
    ma_.close()


def test_argmax_over_more_than_256_queries_matches_numpy(eng):
    """The fused kernel holds 256 queries; the reference has no limit on the number of categories (vlmap.py:123), so a
    wider batch takes the exact scores and the first maximum per row -- from host and from device arrays."""
    import torch

    feat, q = synth.index_inputs(5_000, 64, 300, seed=21)
    q[177] = q[12]                                   # an exact tie across the 256 boundary: the lower index wins
    ref = O.argmax(O.scores(feat, q))
    m = eng.DeviceMap(feat)
    assert np.array_equal(m.argmax(q), ref)
    got = m.argmax(torch.from_numpy(q).cuda())
    assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), ref)
    m.close()


def test_pipelined_topk_calls_return_the_same_results(eng):
    """AVL_PIPELINED: the tail of a call (finalize beside the next call's screen, fallback) runs on the map's tail stream.
    A run of pipelined calls with different query batches, ring buffers and a flush must return what the plain calls
    return -- including a batch whose queries all overflow into the device-side fallback."""
    import torch

    feat, _ = synth.index_inputs(60_000, 512, 1, seed=12)
    feat[40_000:] = feat[7]                              # 20 000 copies of one row: queries near it overflow
    m = eng.DeviceMap(feat)
    batches = [synth.index_inputs(1, 512, 256, seed=50 + i)[1] for i in range(5)]
    batches[2][:40] = feat[7] / np.linalg.norm(feat[7])   # 40 queries aligned with the duplicated row -> fallback
    want = [m.topk(b, 16) for b in batches]
    qd = [torch.from_numpy(b).cuda() for b in batches]
    outs = [(torch.empty((256, 16), dtype=torch.int64, device="cuda"), torch.empty((256, 16), dtype=torch.float32, device="cuda"))
            for _ in batches]
    for rep in range(2):
        for q, o in zip(qd, outs):
            m.topk(q, 16, out=o, stats=False, pipelined=True)
        m.flush()
        torch.cuda.synchronize()
        for (wi, wv), (oi, ov) in zip(want, outs):
            assert np.array_equal(oi.cpu().numpy(), wi) and np.array_equal(ov.cpu().numpy(), wv)
    # a plain call right after pipelined ones (it must wait for their tails by itself)
    m.topk(qd[0], 16, out=outs[0], stats=False, pipelined=True)
    i2, v2 = m.topk(batches[1], 16)
    assert np.array_equal(i2, want[1][0]) and np.array_equal(v2, want[1][1])
    m.close()
