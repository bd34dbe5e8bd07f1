"""CPU property tests (hypothesis) of the host-side logic that has no GPU in it: the shard merge, the slab
split, the first-touch key ranking of the sharded build, and the oracle's order-free fusion form."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from avlmaps_b200.sharded import merge_topk, slab_bounds
from oracle import avl_oracle as O


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 400), st.integers(1, 9), st.integers(1, 20), st.integers(0, 2 ** 31 - 1))
def test_merge_of_any_slab_split_equals_global_topk(n, world, k, seed):
    """Per-slab top-k (score desc, row asc) merged == top-k over all rows, for any split, with heavy ties."""
    rng = np.random.default_rng(seed)
    scores = rng.integers(0, 5, n).astype(np.float32)  # ties everywhere
    parts_i, parts_v = [], []
    for r in range(world):
        lo, hi = slab_bounds(n, world, r)
        i, v = O.topk_vector(scores[lo:hi], k) if hi > lo else (np.full(k, -1, np.int64), np.full(k, -np.inf, np.float32))
        parts_i.append(np.where(i >= 0, i + lo, -1)[None])
        parts_v.append(v[None])
    mi, mv = merge_topk(np.stack(parts_i), np.stack(parts_v), k)
    ri, rv = O.topk_vector(scores, k)
    assert np.array_equal(mi[0], ri) and np.array_equal(mv[0], rv)


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 10_000), st.integers(1, 64))
def test_slab_bounds_partition(n, world):
    spans = [slab_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 300), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
def test_key_ranking_recovers_global_first_touch_order(v, world, seed):
    """Sharded build: voxels carry unique first-touch keys; a slab's local ids are its keys in ascending order;
    the rank of a key among all keys is the voxel's id in a single build (what avl_rank_keys computes)."""
    rng = np.random.default_rng(seed)
    keys = np.sort(rng.choice(2 ** 40, v, replace=False).astype(np.uint64))          # global first-touch order
    owner = rng.integers(0, world, v)
    shards = [keys[owner == r] for r in range(world)]
    allk = np.sort(np.concatenate(shards))
    for r in range(world):
        gids = np.searchsorted(allk, shards[r])
        assert np.array_equal(keys[gids], shards[r]) and np.all(np.diff(gids) > 0) if gids.size > 1 else True


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 40), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
def test_closed_form_fusion_matches_sequential_update(n_obs, d, seed):
    """g = (a0^2 f0 + sum_{i>=1} a_i f_i) / sum a_i equals the reference's running mean
    (first touch stores f*alpha with weight alpha) up to float32 rounding."""
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((n_obs, d)).astype(np.float32)
    a = np.exp(-rng.uniform(0.1, 6.0, n_obs) ** 2 / 1.2)
    g = (f[0].astype(np.float64) * a[0]).astype(np.float32)      # vlmap_builder.py:166
    w = np.float32(a[0])
    for i in range(1, n_obs):                                      # :172-178
        g = ((g * w + f[i].astype(np.float64) * a[i]) / (np.float64(w) + a[i])).astype(np.float32)
        w = np.float32(np.float64(w) + a[i])
    closed = (a[0] ** 2 * f[0].astype(np.float64) + (a[1:, None] * f[1:].astype(np.float64)).sum(0)) / a.sum()
    assert np.allclose(g, closed, rtol=2e-4, atol=1e-6) and np.isclose(w, a.sum(), rtol=1e-5)


_DTYPES = ["<f4", "<f8", "<f2", "<i4", "<i8", "<u1", "<u2", ">i4", ">f8", "<i1", "<u8"]


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(st.sampled_from(_DTYPES), st.lists(st.integers(0, 7), min_size=0, max_size=3)), min_size=1, max_size=9),
       st.integers(0, 2 ** 31 - 1))
def test_h5lite_roundtrip_of_arbitrary_datasets(specs, seed):
    """Any set of numeric datasets -- scalars, empty dimensions, either byte order -- comes back with the same bytes,
    shape and dtype; names are found whatever their sort order in the symbol table."""
    import tempfile
    from pathlib import Path

    from avlmaps_b200.utils import h5lite

    rng = np.random.default_rng(seed)
    data = {}
    for i, (dt, shape) in enumerate(specs):
        a = (rng.standard_normal(shape) * 100).astype(np.dtype(dt)) if np.dtype(dt).kind == "f" else \
            rng.integers(0, 100, shape).astype(np.dtype(dt))
        data[f"{'zyxab'[i % 5]}_{i}_{'grid_feat' if i % 2 else 'w'}"] = a
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "m.h5df"
        h5lite.write_file(p, data)
        with h5lite.File(p) as f:
            assert sorted(f.keys()) == sorted(data)
            for k, v in data.items():
                got = f[k].read()
                assert got.shape == v.shape and got.dtype == v.dtype and got.tobytes() == v.tobytes(), k


@settings(max_examples=50, deadline=None)
@given(st.dictionaries(st.sampled_from(["a", "b", "c", "d"]), st.one_of(st.integers(-5, 5), st.booleans(), st.floats(-2, 2, allow_nan=False),
                                                                        st.sampled_from(["x", "y z", ""])), min_size=1),
       st.sampled_from(["a", "b", "c", "d"]))
def test_config_interpolation_returns_the_target_value_and_type(values, ref_key):
    """`${group.key}` as a whole value yields the referenced value with its type; embedded in text it is formatted."""
    import tempfile
    from pathlib import Path

    import yaml

    from avlmaps_b200.config import ConfigError, compose

    with tempfile.TemporaryDirectory() as td:
        root = Path(td)
        (root / "g").mkdir()
        (root / "g" / "v.yaml").write_text(yaml.safe_dump(values))
        (root / "main.yaml").write_text("defaults:\n  - g: v\n  - _self_\nwhole: ${g." + ref_key + "}\ntext: \"<${g." + ref_key + "}>\"\n")
        if ref_key not in values:
            try:
                compose(root, "main")
                raise AssertionError("a dangling interpolation must fail at compose time")
            except ConfigError:
                return
        c = compose(root, "main")
        assert c.whole == values[ref_key] and type(c.whole) is type(values[ref_key])
        assert c.text == f"<{values[ref_key]}>"
