"""GPU bring-up of the index path, one case per subprocess so a trap/hang in one variant
does not hide the others.  Usage on the GPU box:

    python tools/bringup_index.py            # runs every case under `timeout`, writes gpurun_out/bringup_*.json
    python tools/bringup_index.py --case X   # one case in this process
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"


def bf16_round(x: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32)


def gen(n, d, q, seed=0):
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((n, d), dtype=np.float32)
    s = (14.2857 * np.random.default_rng(seed + 1).uniform(0.05, 1.0, n)).astype(np.float32)
    feat *= s[:, None]
    qq = np.random.default_rng(seed + 2).standard_normal((q, d)).astype(np.float32)
    qq /= np.linalg.norm(qq, axis=1, keepdims=True)
    return feat, qq.astype(np.float32)


def exact_scores(feat, q):
    return (feat.astype(np.float64) @ q.astype(np.float64).T).astype(np.float32)


def case_dense(lib, L, n, d, nq):
    feat, q = gen(n, d, nq)
    m = C.c_void_p()
    L.check(lib.avl_map_create(L.np_ptr(feat), n, d, 0, None, C.byref(m)))
    out = np.empty((n, nq), np.float32)
    L.check(lib.avl_sim_dense(m, L.np_ptr(q), nq, None, 0, L.np_ptr(out), 0, None))
    ref = exact_scores(feat, q)
    lib.avl_map_destroy(m)
    return {"max_abs_diff": float(np.abs(out - ref).max()), "n_bitdiff": int((out != ref).sum()),
            "ok": bool(np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max())}


def case_screen(lib, L, n, d, nq, cg):
    feat, q = gen(n, d, nq)
    m = C.c_void_p()
    L.check(lib.avl_map_create(L.np_ptr(feat), n, d, 0, None, C.byref(m)))
    out = np.full((n, nq), np.nan, np.float32)
    L.check(lib.avl_sim_screen_dense(m, L.np_ptr(q), nq, cg, L.np_ptr(out), 0, None))
    ref = bf16_round(feat).astype(np.float64) @ bf16_round(q).astype(np.float64).T
    lib.avl_map_destroy(m)
    scale = np.linalg.norm(feat, axis=1)[:, None] * np.linalg.norm(q, axis=1)[None, :]
    err = np.abs(out - ref) / scale
    bad = np.argwhere(~(err < 1e-4))
    return {"max_rel_err_vs_bf16_ref": float(np.nanmax(err)), "n_nan": int(np.isnan(out).sum()),
            "n_bad": int(len(bad)), "first_bad": bad[:8].tolist(),
            "sample_out": out[:2, :4].tolist(), "sample_ref": ref[:2, :4].tolist(),
            "ok": bool(len(bad) == 0)}


def case_argmax(lib, L, n, d, nq, cg_env=None, normalize=0):
    if cg_env:
        os.environ["AVL_CTA_GROUP"] = str(cg_env)
    feat, q = gen(n, d, nq)
    m = C.c_void_p()
    L.check(lib.avl_map_create(L.np_ptr(feat), n, d, 0, None, C.byref(m)))
    out = np.full(n, -7, np.int32)
    st = L.IndexStats()
    L.check(lib.avl_sim_argmax(m, L.np_ptr(q), nq, None, normalize, L.np_ptr(out), 0, None, C.byref(st)))
    ref = exact_scores(feat, q)
    ra = ref.argmax(1).astype(np.int32)
    lib.avl_map_destroy(m)
    mism = np.nonzero(out != ra)[0]
    return {"n_mismatch": int(len(mism)), "first": mism[:8].tolist(), "stats": st.as_dict(),
            "flag_frac": st.n_flagged / max(n, 1), "ok": bool(len(mism) == 0)}


def ref_topk(scores, k):
    # (score desc, index asc)
    n = scores.shape[0]
    out_i = np.full((scores.shape[1], k), -1, np.int64)
    out_s = np.full((scores.shape[1], k), -np.inf, np.float32)
    for j in range(scores.shape[1]):
        col = scores[:, j]
        order = np.lexsort((np.arange(n), -col.astype(np.float64)))[:k]
        out_i[j, :len(order)] = order
        out_s[j, :len(order)] = col[order]
    return out_i, out_s


def case_topk(lib, L, n, d, nq, k, cg_env=None, normalize=0, use_scale=False):
    if cg_env:
        os.environ["AVL_CTA_GROUP"] = str(cg_env)
    feat, q = gen(n, d, nq)
    scale = None
    if use_scale:
        scale = np.random.default_rng(5).uniform(0.5, 100.0, nq).astype(np.float32)
    m = C.c_void_p()
    L.check(lib.avl_map_create(L.np_ptr(feat), n, d, 0, None, C.byref(m)))
    oi = np.full((nq, k), -9, np.int64)
    os_ = np.full((nq, k), np.nan, np.float32)
    st = L.IndexStats()
    L.check(lib.avl_sim_topk(m, L.np_ptr(q), nq, L.np_ptr(scale), normalize, k, L.np_ptr(oi), L.np_ptr(os_), 0,
                             None, C.byref(st)))
    ref = exact_scores(feat, q)
    if normalize:
        nrm = np.sqrt((feat.astype(np.float64) ** 2).sum(1)).astype(np.float32)
        inv = np.where(nrm > 0, np.float32(1.0) / nrm, np.float32(0)).astype(np.float32)
        ref = ref * inv[:, None]
    if use_scale:
        ref = ref * scale[None, :]
    ri, rs = ref_topk(ref, k)
    lib.avl_map_destroy(m)
    return {"idx_mismatch": int((oi != ri).sum()), "max_score_diff": float(np.nanmax(np.abs(os_ - rs))),
            "stats": st.as_dict(), "ok": bool((oi == ri).all())}


def case_vec_topk(lib, L, n, k):
    rng = np.random.default_rng(3)
    v = rng.standard_normal(n).astype(np.float32)
    v[rng.integers(0, n, n // 3)] = 1.0  # many exact ties at the top
    oi = np.empty(k, np.int64)
    ov = np.empty(k, np.float32)
    L.check(lib.avl_topk_f32(L.np_ptr(v), n, k, L.np_ptr(oi), L.np_ptr(ov), 0, None))
    order = np.lexsort((np.arange(n), -v.astype(np.float64)))[:k]
    return {"ok": bool((oi[:len(order)] == order).all()), "got": oi[:8].tolist(), "want": order[:8].tolist()}


def case_perf(lib, L, n, d, nq, k, mode, cg_env=None, iters=5):
    import torch
    if cg_env:
        os.environ["AVL_CTA_GROUP"] = str(cg_env)
    g = torch.Generator(device="cuda").manual_seed(0)
    feat = torch.randn((n, d), device="cuda", generator=g) * 3.0
    q = torch.randn((nq, d), device="cuda", generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    m = C.c_void_p()
    L.check(lib.avl_map_create(C.c_void_p(feat.data_ptr()), n, d, L.AVL_ON_DEVICE, None, C.byref(m)))
    del feat
    lib.avl_set_profiling(1)
    st = L.IndexStats()
    res = []
    oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    osc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    oa = torch.empty(n, dtype=torch.int32, device="cuda")
    for it in range(iters):
        t0 = time.perf_counter()
        if mode == "topk":
            L.check(lib.avl_sim_topk(m, C.c_void_p(q.data_ptr()), nq, None, 0, k, C.c_void_p(oi.data_ptr()),
                                     C.c_void_p(osc.data_ptr()), L.AVL_ON_DEVICE, None, C.byref(st)))
        else:
            L.check(lib.avl_sim_argmax(m, C.c_void_p(q.data_ptr()), nq, None, 0, C.c_void_p(oa.data_ptr()),
                                       L.AVL_ON_DEVICE, None, C.byref(st)))
        torch.cuda.synchronize()
        res.append({"wall_ms": (time.perf_counter() - t0) * 1e3, **st.as_dict()})
    lib.avl_map_destroy(m)
    best = min(r["ms_screen"] for r in res[1:])
    flops = 2.0 * n * d * nq
    byts = n * d * 2.0
    return {"iters": res, "best_ms_screen": best, "tflops": flops / best / 1e9, "gbs": byts / best / 1e6, "ok": True}


CASES = {
    "dense_small": lambda lib, L: case_dense(lib, L, 1000, 512, 9),
    "dense_odd_d": lambda lib, L: case_dense(lib, L, 777, 100, 3),
    "vec_topk": lambda lib, L: case_vec_topk(lib, L, 1_000_003, 16),
    "screen_cg1_q16": lambda lib, L: case_screen(lib, L, 1000, 512, 16, 1),
    "screen_cg1_q64_multi": lambda lib, L: case_screen(lib, L, 70_000, 512, 64, 1),
    "screen_cg1_q9_d768": lambda lib, L: case_screen(lib, L, 5000, 768, 9, 1),
    "screen_cg2_q16": lambda lib, L: case_screen(lib, L, 1000, 512, 16, 2),
    "screen_cg2_q256": lambda lib, L: case_screen(lib, L, 70_000, 512, 256, 2),
    "screen_cg2_q65_d1024": lambda lib, L: case_screen(lib, L, 9000, 1024, 65, 2),
    "argmax_cg1_q2": lambda lib, L: case_argmax(lib, L, 10_000, 512, 2, 1),
    "argmax_cg1_q64": lambda lib, L: case_argmax(lib, L, 200_000, 512, 64, 1),
    "argmax_cg2_q64": lambda lib, L: case_argmax(lib, L, 200_000, 512, 64, 2),
    "argmax_cg2_q256": lambda lib, L: case_argmax(lib, L, 100_000, 512, 256, 2),
    "topk_cg1_q64_k16": lambda lib, L: case_topk(lib, L, 300_000, 512, 64, 16, 1),
    "topk_cg1_norm_scale": lambda lib, L: case_topk(lib, L, 100_000, 512, 33, 5, 1, normalize=1, use_scale=True),
    "topk_cg2_q256_k16": lambda lib, L: case_topk(lib, L, 300_000, 512, 256, 16, 2),
    "topk_tiny": lambda lib, L: case_topk(lib, L, 50, 512, 3, 16, 1),
    "perf_topk_cg2_4m": lambda lib, L: case_perf(lib, L, 4_194_304, 512, 256, 16, "topk", 2),
    "perf_argmax_cg1_1m_q64": lambda lib, L: case_perf(lib, L, 1_000_000, 512, 64, 16, "argmax", 1),
    "perf_argmax_cg2_1m_q64": lambda lib, L: case_perf(lib, L, 1_000_000, 512, 64, 16, "argmax", 2),
    "perf_topk_cg1_1m_q64": lambda lib, L: case_perf(lib, L, 1_000_000, 512, 64, 16, "topk", 1),
}


def run_case(name):
    from avlmaps_b200 import _lib as L
    lib = L.load()
    L.require_device()
    t0 = time.time()
    try:
        r = CASES[name](lib, L)
    except Exception as e:  # noqa: BLE001
        r = {"ok": False, "error": repr(e)}
    r["case"] = name
    r["seconds"] = time.time() - t0
    print(json.dumps(r))
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=150)
    a = ap.parse_args()
    if a.case:
        run_case(a.case)
        return
    OUT.mkdir(exist_ok=True)
    summary = []
    for name in CASES:
        if a.only and not any(tok in name for tok in a.only.split(",")):
            continue
        p = subprocess.run(["timeout", str(a.timeout), sys.executable, __file__, "--case", name],
                           capture_output=True, text=True)
        line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
        try:
            r = json.loads(line)
        except Exception:  # noqa: BLE001
            r = {"case": name, "ok": False, "rc": p.returncode, "stdout": p.stdout[-2000:], "stderr": p.stderr[-3000:]}
        r["rc"] = p.returncode
        summary.append(r)
        print(("PASS " if r.get("ok") else "FAIL ") + name + " " + json.dumps({k: v for k, v in r.items() if k not in ("iters",)})[:600], flush=True)
        (OUT / "bringup_index.json").write_text(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
