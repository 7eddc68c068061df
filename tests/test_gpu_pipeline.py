"""-m gpu: the frame-render pipeline (localrf_b200/pipeline.py): frames rendered through the ring of pinned
buffers equal direct LocalTensorfs.forward calls; the device-side 8-bit conversions equal the host-side
conversions renderer.py applies (cv2.imwrite's saturating round of 255 * rgb[..., ::-1]; visualize_depth =
cv2.applyColorMap of the truncated, clipped, normalised depth, utils/utils.py:179-197)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_frames_match_direct_calls_and_host_conversions():
    import cv2
    import bench
    import localrf_b200 as L
    wl = bench.Workload("cfg5", 48)                       # 8 small fields, 64 frames, cross-fades
    lt = wl.build(L.LocalTensorfs, quiet=True).to("cuda")
    W, H = 96, 80
    pipe = L.FramePipeline(lt, W, H, depth_minmax=(0.0, 5.0), n_buffers=3, keep_float=True)
    frames = [torch.tensor([f], device="cuda") for f in (0, 5, 6, 7, 8, 30, 63)]
    # the yielded arrays are views of the ring's pinned buffers (valid until n_buffers more submits): copy
    got = [{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in fr.items()}
           for fr in pipe.render(frames, floater_thresh=0.5)]
    assert len(got) == len(frames) and pipe.in_flight() == 0
    ids = torch.arange(W * H, dtype=torch.int64, device="cuda")
    for view, fr in zip(frames, got):
        with torch.no_grad():
            rgb, depth, _, _ = lt(ids, view, W, H, is_train=False, floater_thresh=0.5)
        rgb, depth = rgb.cpu().numpy().reshape(H, W, 3), depth.cpu().numpy().reshape(H, W)
        np.testing.assert_array_equal(fr["rgb"], rgb)                     # zero-copy float outputs, bit-exact
        np.testing.assert_array_equal(fr["depth"], depth)
        want_rgb8 = np.clip(np.rint(255.0 * rgb[..., ::-1].astype(np.float32)), 0, 255).astype(np.uint8)
        np.testing.assert_array_equal(fr["rgb8"], want_rgb8)
        x = (np.nan_to_num(depth) - 0.0) / (5.0 - 0.0 + 1e-8)
        idx = (255 * np.clip(x, 0, 1)).astype(np.uint8)
        want_d8 = cv2.applyColorMap(idx, cv2.COLORMAP_JET)
        mism = (fr["depth8"] != want_d8).any(-1)
        # numpy evaluates (d - lo) / (hi - lo + 1e-8) in float64 (python scalars), the kernel in fp32: an index
        # may differ by one where 255 * x sits on an integer to rounding
        assert mism.mean() < 2e-3, float(mism.mean())


def test_pipeline_refuses_overflow_and_cpu_model():
    import bench
    import localrf_b200 as L
    lt = bench.build_scene("cuda", 32)
    pipe = L.FramePipeline(lt, 32, 24, n_buffers=2)
    v = torch.tensor([0], device="cuda")
    pipe.submit(v); pipe.submit(v)
    with pytest.raises(RuntimeError, match="in flight"):
        pipe.submit(v)
    assert pipe.fetch() is not None and pipe.fetch() is not None and pipe.fetch() is None
    with pytest.raises(RuntimeError, match="CUDA"):
        L.FramePipeline(bench.build_scene("cpu", 16), 8, 8)
