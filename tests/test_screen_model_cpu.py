"""CPU: a numpy MODEL of the index path's screening logic, checked against the oracle on
adversarial inputs.  It restates the formulas of the kernels -- per-row / per-query statistics
(csrc/sim_exact.cu map_prepare_kernel, query_prepare_kernel), the band and the threshold epilogue
(csrc/sim_screen.cu), the finalize step (topk_finalize_kernel), the argmax margin test -- with the tensor-core product
replaced by a float32 matmul of the rounded operands, and verifies the two claims the design rests on:

  1. the band is rigorous: |S~ - s| <= (||a - a~|| + kappa ||a~|| + rho ||a~||) ||b||  for every (row, query);
  2. screening by bounds never loses an answer: (threshold from lower bounds of a strided sample) + (candidates by
     upper bound) + (exact re-score of the survivors) returns exactly the oracle's top-k, and (margin test) + (exact
     re-rank of the in-band queries) returns exactly the oracle's argmax -- with ties, duplicates, wide norm ranges.

It runs no product code (that needs a B200; the GPU suite tests the kernels themselves against the same oracle)."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import avl_oracle as O

F32 = np.float32


def round_operand(x: np.ndarray, f16: bool) -> np.ndarray:
    """bf16 (round to nearest even on the upper 16 bits) or fp16 rounding of float32 values, returned as float32."""
    x = np.ascontiguousarray(x, F32)
    if f16:
        return x.astype(np.float16).astype(F32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(F32)


def up(x):
    """float32 rounded towards +inf from a float64 value (the kernels' __double2float_ru)."""
    y = np.asarray(x, np.float64).astype(F32)
    return np.where(y.astype(np.float64) < x, np.nextafter(y, F32(np.inf)), y).astype(F32)


class ScreenModel:
    def __init__(self, feat, q, f16=False):
        self.feat, self.q = feat.astype(F32), q.astype(F32)
        d = feat.shape[1]
        dpad = (d + 63) // 64 * 64
        kappa = F32(dpad) * F32(2.4e-7)
        self.at, self.bt = round_operand(self.feat, f16), round_operand(self.q, f16)
        a64, at64 = self.feat.astype(np.float64), self.at.astype(np.float64)
        self.row_norm = np.sqrt((a64 * a64).sum(1)).astype(F32)
        self.row_an = up(np.sqrt((at64 * at64).sum(1)) * (1.0 + 1e-7))
        self.row_c = up(np.sqrt(((a64 - at64) ** 2).sum(1)) * (1.0 + 1e-7) + np.float64(kappa) * self.row_an.astype(np.float64))
        b64, bt64 = self.q.astype(np.float64), self.bt.astype(np.float64)
        bn = np.sqrt((b64 * b64).sum(1))
        self.q_bn = up(bn * (1.0 + 1e-7))
        self.rho = F32(up(np.max(np.sqrt(((b64 - bt64) ** 2).sum(1)) / np.maximum(bn, 1e-300)) * (1.0 + 1e-6) + 2e-6))
        self.s_tilde = (self.at @ self.bt.T).astype(F32)          # stand-in for the tcgen05 product (fp32 accumulation)
        self.r = (self.rho * self.row_an + self.row_c).astype(F32)  # fmaf(rho, row_an, row_c)

    def band(self):
        return self.r.astype(np.float64)[:, None] * self.q_bn.astype(np.float64)[None, :] * (1 + 1e-6)

    def topk(self, k, normalize=False, group=32, stride=4):
        n, nq = self.s_tilde.shape
        w = np.maximum(self.row_norm, F32(1e-30)).astype(np.float64) if normalize else np.ones(n)
        eps = self.band()
        lb = (self.s_tilde.astype(np.float64) - eps) / w[:, None]
        lb = lb - np.abs(lb) * 2.0 ** -20                        # the epilogue pushes the rounding down
        ub = (self.s_tilde.astype(np.float64) + eps) / w[:, None]
        ub = ub + np.abs(ub) * 2.0 ** -20
        exact = O.scores(self.feat, self.q, normalize=normalize)
        out_i = np.full((nq, k), -1, np.int64)
        out_v = np.full((nq, k), -np.inf, F32)
        n_cand = 0
        for j in range(nq):
            # threshold: k-th largest of the maxima of disjoint 32-row groups of a strided sample (select_threshold)
            rows = np.arange(0, n, stride)
            gmax = [lb[rows[g:g + group], j].max() for g in range(0, rows.size, group)]
            t = np.sort(gmax)[-k] if len(gmax) >= k else -np.inf
            cand = np.nonzero(ub[:, j] >= t)[0]
            n_cand += cand.size
            # finalize: survivors reach the k-th best lower bound among the candidates; exact re-score; (score desc, row asc)
            kk = min(k, cand.size)
            v = np.sort(lb[cand, j])[-kk] if kk else -np.inf
            surv = cand[ub[cand, j] >= v]
            order = np.lexsort((surv, -exact[surv, j].astype(np.float64)))[:k]
            out_i[j, :order.size] = surv[order]
            out_v[j, :order.size] = exact[surv[order], j]
        return out_i, out_v, n_cand

    def argmax(self, normalize=False):
        """Margin test of the argmax epilogue: decided by the screen unless the top-2 margin is within 2 eps_i."""
        exact = O.scores(self.feat, self.q, normalize=normalize)
        s = self.s_tilde.astype(np.float64)
        bn_max = float(self.q_bn.max())
        tol = (2.0 * self.r.astype(np.float64) + 6.2e-5 * self.row_an.astype(np.float64)) * bn_max * 1.0001
        best = s.max(1)
        out = s.argmax(1).astype(np.int32)
        flagged = 0
        for i in range(s.shape[0]):
            inband = np.nonzero(s[i] >= best[i] - tol[i])[0]
            if inband.size > 1:
                flagged += 1
                out[i] = inband[np.argmax(exact[i, inband])]       # first maximum among the in-band queries
        return out, flagged


def lseg_like(n, d, nq, seed):
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((n, d)).astype(F32) * (14.2857 * rng.uniform(0.05, 1, n)).astype(F32)[:, None] / np.sqrt(d).astype(F32)
    q = rng.standard_normal((nq, d)).astype(F32)
    return feat, (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(F32)


def adversarial(n, d, nq, seed):
    """Norms over ten orders of magnitude, exact duplicates (ties across rows), near-duplicates one ulp apart, a zero
    row, rows aligned with a query (scores near the maximum), cancelling components."""
    rng = np.random.default_rng(seed)
    feat, q = lseg_like(n, d, nq, seed)
    feat *= (10.0 ** rng.uniform(-5, 5, n)).astype(F32)[:, None]
    feat[n // 3] = feat[5]
    feat[n // 2] = feat[5]
    feat[7] = np.nextafter(feat[5], F32(np.inf))
    feat[11] = 0
    feat[13] = q[0] * F32(1e5)
    feat[17] = q[0] * F32(1e5) * F32(1 + 2 ** -10)
    feat[19, ::2] = F32(3e4)
    feat[19, 1::2] = F32(-3e4)
    return feat, q


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("gen,n,d,nq", [(lseg_like, 3000, 512, 9), (adversarial, 2000, 100, 5), (adversarial, 1500, 64, 33)])
def test_error_band_is_rigorous(gen, n, d, nq, f16):
    feat, q = gen(n, d, nq, seed=1)
    if f16 and np.abs(feat).max() > 6e4:
        feat = feat * F32(6e4 / np.abs(feat).max())            # the library falls back to bf16 beyond the fp16 range
    m = ScreenModel(feat, q, f16)
    exact = feat.astype(np.float64) @ q.astype(np.float64).T
    err = np.abs(m.s_tilde.astype(np.float64) - exact)
    assert np.all(err <= m.band())
    # and it is not vacuous: the band is within a few hundred times the observed error where that error is largest
    i, j = np.unravel_index(np.argmax(err / np.maximum(m.band(), 1e-300)), err.shape)
    assert err[i, j] >= m.band()[i, j] / 500


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("gen,n,d,nq,k", [(lseg_like, 4000, 512, 8, 16), (adversarial, 3000, 100, 6, 5), (adversarial, 900, 64, 17, 32)])
def test_screened_topk_equals_the_oracle(gen, n, d, nq, k, normalize, f16):
    feat, q = gen(n, d, nq, seed=2)
    if f16 and np.abs(feat).max() > 6e4:
        feat = feat * F32(6e4 / np.abs(feat).max())
    if normalize:
        feat[np.linalg.norm(feat, axis=1) == 0] = F32(1e-3)    # 0 / 0 is NaN in the reference too: not a case
    m = ScreenModel(feat, q, f16)
    idx, val, n_cand = m.topk(k, normalize=normalize)
    ri, rv = O.topk(O.scores(feat, q, normalize=normalize), k)
    assert np.array_equal(idx, ri) and np.array_equal(val, rv)
    if -(-n // 4) // 32 >= k:                                   # enough sample groups for a finite threshold
        assert n_cand < 0.5 * n * nq                            # ... then the screen does screen
    else:
        assert n_cand == n * nq                                 # a tiny map: threshold -inf, everything is re-scored


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("gen,n,d,nq", [(lseg_like, 3000, 512, 2), (lseg_like, 2000, 512, 64), (adversarial, 2500, 100, 9)])
def test_margin_test_plus_rerank_equals_the_oracle_argmax(gen, n, d, nq, f16):
    feat, q = gen(n, d, nq, seed=3)
    if f16 and np.abs(feat).max() > 6e4:
        feat = feat * F32(6e4 / np.abs(feat).max())
    m = ScreenModel(feat, q, f16)
    got, flagged = m.argmax()
    assert np.array_equal(got, O.argmax(O.scores(feat, q)))
    assert flagged < feat.shape[0]                              # some rows are decided by the screen alone
