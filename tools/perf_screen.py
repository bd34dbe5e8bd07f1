"""A/B harness of the headline step (4 194 304 x 512 map, 256 queries, top-16) on one GPU.

    python tools/perf_screen.py --variants "base:" "pf2:AVL_PREFETCH_TILES=2" "rowmajor:AVL_TILED=0" --steps 20 200

Every variant is `name:ENV=VAL,ENV=VAL`; each (variant, steps) pair runs in its own process (the switches are read once
per process), builds the map on the device, runs 5 warm-up calls and `steps` timed calls through the C-ABI with device
pointers, and reports the library's own CUDA-event timings (`ms_screen` = the main screen kernel, `ms_total` = the
whole call), the wall time per call, the candidate count and the SM clock sampled DURING the timed loop (pynvml, 2 ms
period) -- 20 steps stay at the burst clock, 200+ steps show the power-capped steady state."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


class ClockSampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.samples, self.power, self.stop = [], [], False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.002)


def child(steps: int, n: int, d: int, nq: int, k: int, mode: str):
    import torch

    from avlmaps_b200 import _lib as L

    lib = L.load()
    L.require_device()
    g = torch.Generator(device="cuda").manual_seed(0)
    feat = torch.randn((n, d), device="cuda", generator=g) * 3.0
    q = torch.randn((nq, d), device="cuda", generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    m = C.c_void_p()
    flags = L.AVL_ON_DEVICE | (L.AVL_MAP_F16 if os.environ.get("PERF_F16") == "1" else 0)
    L.check(lib.avl_map_create(C.c_void_p(feat.data_ptr()), n, d, flags, None, C.byref(m)))
    del feat
    lib.avl_set_profiling(1)
    st = L.IndexStats()
    oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    osc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    oa = torch.empty(n, dtype=torch.int32, device="cuda")

    def call():
        if mode == "topk":
            L.check(lib.avl_sim_topk(m, C.c_void_p(q.data_ptr()), nq, None, 0, k, C.c_void_p(oi.data_ptr()),
                                     C.c_void_p(osc.data_ptr()), L.AVL_ON_DEVICE, None, C.byref(st)))
        else:
            L.check(lib.avl_sim_argmax(m, C.c_void_p(q.data_ptr()), nq, None, 0, C.c_void_p(oa.data_ptr()),
                                       L.AVL_ON_DEVICE, None, C.byref(st)))

    for _ in range(5):
        call()
    torch.cuda.synchronize()
    time.sleep(0.5)            # let the clocks recover: every variant starts from the same state
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    smp = ClockSampler()
    smp.start()
    scr, tot = [], []
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
        scr.append(st.ms_screen)
        tot.append(st.ms_total)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    smp.stop = True
    smp.join(timeout=1)
    chk = int(oi.sum().item()) if mode == "topk" else int(oa.sum().item())
    # the same calls enqueued asynchronously (device pointers, no stats): the pipelined step the bench's `value` times
    async_ms = async_scr = pipe_ms = pipe_scr = None
    if mode == "topk":
        time.sleep(0.5)
        for _ in range(3):
            L.check(lib.avl_sim_topk(m, C.c_void_p(q.data_ptr()), nq, None, 0, k, C.c_void_p(oi.data_ptr()),
                                     C.c_void_p(osc.data_ptr()), L.AVL_ON_DEVICE, None, None))
        torch.cuda.synchronize()
        buf = (C.c_float * 256)()
        nn = C.c_int32(0)
        L.check(lib.avl_map_screen_times(m, buf, 256, C.byref(nn)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            L.check(lib.avl_sim_topk(m, C.c_void_p(q.data_ptr()), nq, None, 0, k, C.c_void_p(oi.data_ptr()),
                                     C.c_void_p(osc.data_ptr()), L.AVL_ON_DEVICE, None, None))
        e1.record()
        torch.cuda.synchronize()
        async_ms = e0.elapsed_time(e1) / steps
        L.check(lib.avl_map_screen_times(m, buf, 256, C.byref(nn)))
        async_scr = statistics.median(buf[i] for i in range(nn.value)) if nn.value else None
        assert int(oi.sum().item()) == chk, "asynchronous calls returned a different result"
        # ... and pipelined: the tail of a call runs next to the following call's screen
        outs = [(torch.empty_like(oi), torch.empty_like(osc)) for _ in range(4)]
        for _ in range(3):
            L.check(lib.avl_sim_topk(m, C.c_void_p(q.data_ptr()), nq, None, 0, k, C.c_void_p(outs[0][0].data_ptr()),
                                     C.c_void_p(outs[0][1].data_ptr()), L.AVL_ON_DEVICE | L.AVL_PIPELINED, None, None))
        L.check(lib.avl_map_flush(m, None))
        torch.cuda.synchronize()
        L.check(lib.avl_map_screen_times(m, buf, 256, C.byref(nn)))
        e0.record()
        for i in range(steps):
            L.check(lib.avl_sim_topk(m, C.c_void_p(q.data_ptr()), nq, None, 0, k, C.c_void_p(outs[i % 4][0].data_ptr()),
                                     C.c_void_p(outs[i % 4][1].data_ptr()), L.AVL_ON_DEVICE | L.AVL_PIPELINED, None, None))
        L.check(lib.avl_map_flush(m, None))
        e1.record()
        torch.cuda.synchronize()
        pipe_ms = e0.elapsed_time(e1) / steps
        L.check(lib.avl_map_screen_times(m, buf, 256, C.byref(nn)))
        pipe_scr = statistics.median(buf[i] for i in range(nn.value)) if nn.value else None
        for o in outs[:min(4, steps)]:
            assert int(o[0].sum().item()) == chk, "pipelined calls returned a different result"
    out = {"steps": steps, "ms_screen_min": min(scr), "ms_screen_med": statistics.median(scr),
           "ms_total_med": statistics.median(tot), "wall_ms": wall, "cands": int(st.n_candidates),
           "flagged": int(st.n_flagged), "fallback": int(st.n_fallback_queries), "cg": int(st.cta_group), "checksum": chk,
           "sm_mhz_med": statistics.median(smp.samples) if smp.samples else None,
           "sm_mhz_min": min(smp.samples) if smp.samples else None,
           "power_w_max": max(smp.power) if smp.power else None, "async_step_ms": async_ms, "async_screen_med": async_scr,
           "pipelined_step_ms": pipe_ms, "pipelined_screen_med": pipe_scr}
    lib.avl_map_destroy(m)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", nargs="*", default=["base:"])
    ap.add_argument("--steps", nargs="*", type=int, default=[20])
    ap.add_argument("--n", type=int, default=4_194_304)
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--nq", type=int, default=256)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--mode", default="topk")
    ap.add_argument("--child", type=int, default=0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    if a.child:
        child(a.child, a.n, a.d, a.nq, a.k, a.mode)
        return
    rows = []
    for v in a.variants:
        name, _, envs = v.partition(":")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            key, _, val = kv.partition("=")
            env[key] = val
        for steps in a.steps:
            cmd = [sys.executable, __file__, "--child", str(steps), "--n", str(a.n), "--d", str(a.d), "--nq", str(a.nq),
                   "--k", str(a.k), "--mode", a.mode]
            p = subprocess.run(["timeout", "300", *cmd], env=env, capture_output=True, text=True)
            try:
                r = json.loads(p.stdout.strip().splitlines()[-1])
            except Exception:  # noqa: BLE001
                r = {"error": (p.stdout[-500:] + p.stderr[-1500:])}
            r.update(variant=name, env=envs)
            rows.append(r)
            if "error" in r:
                print(f"{name:>14s} steps={steps}: FAILED {r['error'][-600:]}", flush=True)
            else:
                print(f"{name:>14s} steps={steps:4d}: screen min {r['ms_screen_min']:.4f} med {r['ms_screen_med']:.4f}  call {r['ms_total_med']:.4f}"
                      f"  wall {r['wall_ms']:.4f}  clk {r['sm_mhz_med']} (min {r['sm_mhz_min']})  P {r['power_w_max']}"
                      f"  cands {r['cands']} fb {r['fallback']} chk {r['checksum']}"
                      f"  | async step {r.get('async_step_ms') or 0:.4f} screen {r.get('async_screen_med') or 0:.4f}"
                      f"  | pipelined step {r.get('pipelined_step_ms') or 0:.4f} screen {r.get('pipelined_screen_med') or 0:.4f}", flush=True)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
