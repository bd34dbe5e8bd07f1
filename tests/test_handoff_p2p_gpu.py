"""fp16 feature hand-off (AVL_FEAT_F16, csrc/build_path.cu chw16_to_hwc_kernel) and the fused peer-memory exchange with a
world of one (csrc/p2p_exchange.cu; N > 1: tools/p2p_check.py and tools/sharded_index_check.py under torchrun).
Written at the end of round 1 without GPU time; first run on a B200 in round 2 (gpurun_out/r2a_call.log: 4 passed)."""
import os

import numpy as np
import pytest

import synth
from oracle import avl_oracle as O

pytestmark = [pytest.mark.gpu]


def test_fp16_features_give_the_same_map_as_their_float32_values(lib):
    """LSeg's output is fp16-exact (lseg_net.py:318-321): a frame handed over as float16 must build the very same map
    as the same values handed over as float32, from host and from device pointers, odd shapes included."""
    import torch

    from avlmaps_b200 import engine

    for (h, w, fh, fw, d) in ((60, 80, 49, 65, 32), (30, 40, 25, 33, 6)):     # P = fh * fw odd / D not a multiple of 4
        cfg = synth.map_config(48, 0.1, 1.6, [40, 0, 40, 0, 40, 30, 0, 0, 1] if h == 60 else [20, 0, 20, 0, 20, 15, 0, 0, 1], 1)
        poses = synth.circle_poses(3, radius=0.3)
        depths, rgbs, feats = synth.build_inputs(3, h, w, fh, fw, d, seed=8)
        feats16 = [f.astype(np.float16) for f in feats]
        feats32 = [f.astype(np.float32) for f in feats16]                     # the fp16-exact values as float32
        np.random.seed(4)
        sidx = [O.sample_order(h * w, 1) for _ in range(3)]
        b2c, bt = O.setup_transforms(cfg["pose_info"])
        tfs = O.frame_transforms(poses, b2c, bt)
        calib = np.array(cfg["cam_calib_mat"], np.float64).reshape(3, 3)
        want = O.build_map(cfg, poses, depths, rgbs, feats32, sidx, capacity=48 * 48 * 16)
        outs = []
        from avlmaps_b200 import _lib as L

        for variant in ("f32", "f16_host", "f16_device", "f16_hwc_device", "f16_hwc_host", "f16_hwc_batched"):
            b = engine.DeviceBuilder(48, 16, 0.1, d)
            frames = []
            for i in range(3):
                f = feats32[i] if variant == "f32" else feats16[i]
                layout = L.FEAT_CHW
                if "hwc" in variant:      # pixel-major fp16 rows: the hand-off of an encoder that stays on the GPU
                    f, layout = np.ascontiguousarray(np.transpose(f[0], (1, 2, 0))), L.FEAT_HWC
                dd, ss, rr = depths[i], sidx[i], rgbs[i]
                if variant.endswith("device") or variant.endswith("batched"):
                    f, dd, ss, rr = (torch.from_numpy(x).cuda() for x in (f, dd, ss, rr))
                kw = dict(depth=dd, feat=f, kinv=np.linalg.inv(calib), k=calib, kfeat=O.get_sim_cam_mat(fh, fw), tf=tfs[i], rgb=rr,
                          sample_idx=ss, feat_layout=layout)
                if variant.endswith("batched"):
                    frames.append(kw)
                else:
                    b.add_frame(kw.pop("depth"), kw.pop("feat"), kw.pop("kinv"), kw.pop("k"), kw.pop("kfeat"), kw.pop("tf"), **kw)
            if frames:
                b.add_frames(frames)      # one geometry / id-scan / scatter launch triple for the three frames
            outs.append(b.export())
            b.close()
        for o in outs:
            assert np.array_equal(o["grid_pos"], want["grid_pos"]) and np.array_equal(o["occupied_ids"], want["occupied_ids"])
            assert np.allclose(o["grid_feat"], want["grid_feat"], rtol=1e-3, atol=1e-5)
        # the three hand-offs feed identical float32 values to the same kernels in the same order
        assert np.array_equal(outs[0]["grid_pos"], outs[1]["grid_pos"]) and np.array_equal(outs[1]["grid_pos"], outs[2]["grid_pos"])
        assert np.allclose(outs[0]["grid_feat"], outs[1]["grid_feat"], rtol=1e-5, atol=1e-6)
        assert np.allclose(outs[1]["grid_feat"], outs[2]["grid_feat"], rtol=1e-5, atol=1e-6)
        for o in outs[3:]:
            assert np.array_equal(outs[0]["grid_pos"], o["grid_pos"]) and np.allclose(outs[0]["grid_feat"], o["grid_feat"], rtol=1e-5, atol=1e-6)


def test_p2p_exchange_with_a_world_of_one(lib):
    """One rank exchanging with itself: stores into its own receive buffer, waits for its own flags, merges -> the
    input re-ordered by (score desc, row asc), -1 slots last; repeated calls alternate the parity buffers."""
    import torch

    from avlmaps_b200 import engine

    ex = engine.P2PExchange()
    rng = np.random.default_rng(0)
    for nq, k in ((256, 16), (7, 128), (1, 1), (64, 5)):
        idx = np.stack([rng.permutation(10_000)[:k] for _ in range(nq)]).astype(np.int64)
        val = rng.integers(0, 6, (nq, k)).astype(np.float32)          # many ties
        idx[:, k // 2:] = np.where(rng.random((nq, k - k // 2)) < 0.3, -1, idx[:, k // 2:])
        val[idx < 0] = -np.inf
        off = 1_000_000 * nq                                           # slab-local rows -> global rows inside the kernel
        oi, ov = ex.exchange_merge(torch.from_numpy(idx).cuda(), torch.from_numpy(val).cuda(), row_offset=off)
        oi, ov = oi.cpu().numpy(), ov.cpu().numpy()
        idx = np.where(idx >= 0, idx + off, idx)
        assert ex.timed_out_source() == -1
        for q in range(nq):
            keep = idx[q] >= 0
            order = np.lexsort((idx[q][keep], -val[q][keep].astype(np.float64)))
            n = order.size
            assert np.array_equal(oi[q, :n], idx[q][keep][order]) and np.array_equal(ov[q, :n], val[q][keep][order])
            assert np.all(oi[q, n:] == -1) and np.all(np.isneginf(ov[q, n:]))
    ex.close()
