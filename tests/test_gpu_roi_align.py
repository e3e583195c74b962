"""GPU parity: RoIAlignRotated v0/v1 forward+backward and the fused multi-level extractor through the
jdet mirror -> C ABI vs the oracle.  Contract: forward 1e-5 relative, backward 1e-4 relative
(relative to the tensor's magnitude: |got-want| <= tol*(|want| + max|want|))."""
import numpy as np
import pytest
import torch

import workloads as W
from helpers import close_report

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check(got, want, tol, what, max_outliers=0):
    scale = float(np.abs(want).max()) or 1.0
    nbad, maxerr, _ = close_report(got, want, tol, tol * scale)
    print(f"{what}: max abs err {maxerr:.3g} (scale {scale:.3g}), violations {nbad}/{want.size}")
    assert nbad <= max_outliers, what


@pytest.mark.parametrize("version", [0, 1])
@pytest.mark.parametrize("C,sr,out", [(256, 2, 7), (8, 2, 7), (64, 2, (3, 5)), (16, 0, 7), (6, 2, 7)])
def test_single_level_fwd_bwd(cuda, oracle, version, C, sr, out):
    from rs_detection_b200.jdet.ops import roi_align_rotated, roi_align_rotated_v1
    rng = np.random.default_rng(C)
    feat = rng.standard_normal((2, C, 48, 40)).astype(np.float32)
    rois = W.proposals(80, C, batch=2, canvas=640)
    rois[:4, 3:5] = [[0.5, 0.5], [3000, 20], [20, 3000], [1, 900]]
    rois[4, 1:3] = [-40, 1100]
    cls = roi_align_rotated_v1.ROIAlignRotated_v1 if version else roi_align_rotated.ROIAlignRotated
    layer = cls(out, 1 / 16., sr)
    osz = layer.output_size
    x = _t(feat).requires_grad_(True)
    y = layer(x, _t(rois))
    want = oracle.roi_align_rotated_fwd(feat, rois, osz, 1 / 16., sr, version)
    _check(y.detach().cpu().numpy(), want, 1e-5, f"fwd v{version} C={C} sr={sr}")
    g = rng.standard_normal(want.shape).astype(np.float32)
    y.backward(_t(g))
    wantb = oracle.roi_align_rotated_bwd(g, rois, feat.shape, 1 / 16., sr, version)
    _check(x.grad.cpu().numpy(), wantb, 1e-4, f"bwd v{version} C={C} sr={sr}")
    assert "output_size=" in repr(layer) and layer.output_size == osz


def test_reference_selftest_shapes(cuda, oracle):
    # roi_align_rotated_v1.py:376-383
    from rs_detection_b200.jdet.ops.roi_align_rotated_v1 import ROIAlignRotated_v1
    feat = np.random.default_rng(0).standard_normal((2, 1024, 64, 64)).astype(np.float32)
    roi = np.array([[0, 20, 120, 80, 195.5, 0.3], [1, 23, 56, 200, 300.5, 0.2]], np.float32)
    y = ROIAlignRotated_v1((7, 7), 1 / 16.)  # sampling_ratio defaults to 0 (adaptive grid)
    out = y(_t(feat), _t(roi))
    assert tuple(out.shape) == (2, 1024, 7, 7)
    _check(out.cpu().numpy(), oracle.roi_align_rotated_fwd(feat, roi, (7, 7), 1 / 16., 0, 1), 1e-5, "selftest")


@pytest.mark.parametrize("K,batch", [(512, 1), (2000, 2)])
def test_fused_extractor_vs_oracle(cuda, oracle, K, batch):
    """OrientedSingleRoIExtractor on a 4-level pyramid (smaller maps than the bench so the CPU oracle
    finishes in seconds): extension (1.4,1.2), finest_scale 56, 7x7, sampling_ratio 2."""
    from rs_detection_b200.jdet.models.roi_extractors.oriented_single_level import OrientedSingleRoIExtractor
    tile, C = 512, 32
    feats = W.fpn_pyramid(batch, 3, tile=tile, channels=C)
    rois = W.proposals(K, 5, batch=batch, canvas=tile)
    ext = OrientedSingleRoIExtractor(dict(type='ROIAlignRotated_v1', output_size=7, sampling_ratio=2), C,
                                     [4, 8, 16, 32], extend_factor=(1.4, 1.2))
    xs = [_t(f).requires_grad_(True) for f in feats]
    y = ext(xs, _t(rois))
    want, lv = oracle.oriented_extractor_fwd(feats, rois, [4, 8, 16, 32])
    assert len(set(lv.tolist())) == 4  # the synthetic proposals exercise every level
    lv_gpu = ext.map_roi_levels(ext.roi_rescale(_t(rois), (1.4, 1.2)), 4).cpu().numpy()
    assert np.array_equal(lv_gpu, lv)
    _check(y.detach().cpu().numpy(), want, 1e-5, f"extractor fwd K={K}")
    g = np.random.default_rng(1).standard_normal(want.shape).astype(np.float32)
    y.backward(_t(g))
    wb = oracle.oriented_extractor_bwd(g, [f.shape for f in feats], rois, [4, 8, 16, 32])
    for l in range(4):
        _check(xs[l].grad.cpu().numpy(), wb[l], 1e-4, f"extractor bwd level {l}")


def test_rbox_extractor_v0_vs_oracle(cuda, oracle):
    """RboxSingleRoIExtractor (python/jdet/models/roi_extractors/rbox_single_level.py): v0 layer, levels on the RoI
    as given, no extension."""
    from rs_detection_b200.jdet.models.roi_extractors.rbox_single_level import RboxSingleRoIExtractor
    tile, C, K = 512, 16, 600
    feats = W.fpn_pyramid(1, 7, tile=tile, channels=C)
    rois = W.proposals(K, 9, canvas=tile)
    ext = RboxSingleRoIExtractor(dict(type='ROIAlignRotated', output_size=7, sampling_ratio=2), C, [4, 8, 16, 32])
    xs = [_t(f).requires_grad_(True) for f in feats]
    y = ext(xs, _t(rois))
    want, lv = oracle.oriented_extractor_fwd(feats, rois, [4, 8, 16, 32], extend_factor=(1.0, 1.0), version=0)
    assert np.array_equal(ext.map_roi_levels(_t(rois), 4).cpu().numpy(), lv)
    _check(y.detach().cpu().numpy(), want, 1e-5, "rbox extractor fwd")
    g = np.random.default_rng(2).standard_normal(want.shape).astype(np.float32)
    y.backward(_t(g))
    wb = oracle.oriented_extractor_bwd(g, [f.shape for f in feats], rois, [4, 8, 16, 32], extend_factor=(1.0, 1.0), version=0)
    for l in range(4):
        _check(xs[l].grad.cpu().numpy(), wb[l], 1e-4, f"rbox extractor bwd level {l}")


def test_single_feature_list_and_empty(cuda, oracle):
    from rs_detection_b200.jdet.models.roi_extractors.oriented_single_level import OrientedSingleRoIExtractor
    from rs_detection_b200 import core
    feat = np.random.default_rng(2).standard_normal((1, 16, 32, 32)).astype(np.float32)
    rois = W.proposals(30, 2, canvas=128)
    ext = OrientedSingleRoIExtractor(dict(type='ROIAlignRotated_v1', output_size=7, sampling_ratio=2), 16, [4],
                                     extend_factor=(1.4, 1.2))
    y = ext([_t(feat)], _t(rois))  # len(feats)==1 branch: no extension (oriented_single_level.py:92-93)
    _check(y.cpu().numpy(), oracle.roi_align_rotated_fwd(feat, rois, (7, 7), 1 / 4, 2, 1), 1e-5, "single")
    cfg = core.make_roi_cfg([feat.shape], [0.25], 7, 2)
    out = core.roi_align_rotated_forward(cfg, [_t(feat)], torch.zeros((0, 6), device="cuda"))
    assert tuple(out.shape) == (0, 16, 7, 7)
    g = core.roi_align_rotated_backward(cfg, torch.zeros((0, 16, 7, 7), device="cuda"), torch.zeros((0, 6), device="cuda"),
                                        [feat.shape])
    assert float(g[0].abs().max()) == 0.0


def test_channels_last_and_transposes(cuda, oracle):
    from rs_detection_b200 import core
    feat = np.random.default_rng(3).standard_normal((2, 40, 33, 29)).astype(np.float32)
    x = _t(feat)
    nhwc = core.nchw_to_nhwc(x)
    assert torch.equal(nhwc, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(core.nhwc_to_nchw(nhwc), x)
    rois = W.proposals(50, 4, batch=2, canvas=400)
    cfg = core.make_roi_cfg([feat.shape], [1 / 16], 7, 2, channels_last=True)
    out = core.roi_align_rotated_forward(cfg, [nhwc], _t(rois))
    _check(out.cpu().numpy(), oracle.roi_align_rotated_fwd(feat, rois, (7, 7), 1 / 16, 2, 1), 1e-5, "channels_last")


def test_full_size_linearity(cuda):
    """BASELINE config-1 sizes (K=2000, C=256, 1024^2 tile): size-independent properties instead of the
    CPU oracle -- the op is linear in the features, and backward is its adjoint:
    <fwd(x), g> == <x, bwd(g)>."""
    from rs_detection_b200 import core
    shapes = W.fpn_shapes()
    gen = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randn(s, device="cuda", generator=gen) for s in shapes]
    zs = [torch.randn(s, device="cuda", generator=gen) for s in shapes]
    rois = _t(W.proposals(2000, 0))
    cfg = core.make_roi_cfg(shapes, [1 / s for s in W.STRIDES], 7, 2, 1, (1.4, 1.2), 56.0)
    fx = core.roi_align_rotated_forward(cfg, xs, rois)
    fz = core.roi_align_rotated_forward(cfg, zs, rois)
    fxz = core.roi_align_rotated_forward(cfg, [2.0 * a - 0.5 * b for a, b in zip(xs, zs)], rois)
    assert torch.allclose(fxz, 2.0 * fx - 0.5 * fz, rtol=1e-4, atol=1e-4)
    g = torch.randn(fx.shape, device="cuda", generator=gen)
    gx = core.roi_align_rotated_backward(cfg, g, rois, shapes)
    lhs = float((fx.double() * g.double()).sum())
    rhs = float(sum((a.double() * b.double()).sum() for a, b in zip(xs, gx)))
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0)


def test_pixel_major_large_patch_and_border(cuda, oracle):
    """The pixel-major forward path (7x7, sampling_ratio 2, C % 256 == 0) on RoIs whose tap patch exceeds the
    4096-pixel dedupe bitmap (long, diagonal: the per-row lists fall back to one pixel per tap), on tiny RoIs
    (all 196 samples inside a few pixels: every pixel feeds many bins) and on RoIs hanging over the border."""
    from rs_detection_b200 import core
    rng = np.random.default_rng(11)
    feat = rng.standard_normal((1, 256, 112, 120)).astype(np.float32)
    rois = W.proposals(64, 21, canvas=480)
    rois[0, 1:] = [240, 220, 430, 12, 0.78]      # ~108 x 3 px at 45 degrees: patch ~ 78 x 78 > 4096
    rois[1, 1:] = [200, 200, 600, 40, -0.70]
    rois[2, 1:] = [100, 100, 2.0, 1.0, 0.3]      # clamped to 1 x 1 feature pixel
    rois[3, 1:] = [101.3, 57.9, 6.0, 9.0, 1.2]
    rois[4, 1:] = [-30, 470, 200, 90, 0.5]       # mostly outside
    rois[5, 1:] = [4000, 4000, 50, 50, 0.0]      # entirely outside: zeros
    rois[6, 1:] = [478, 2, 64, 64, -1.5]
    for version in (1, 0):
        cfg = core.make_roi_cfg([feat.shape], [0.25], 7, 2, version)
        got = core.roi_align_rotated_forward(cfg, [_t(feat)], _t(rois)).cpu().numpy()
        want = oracle.roi_align_rotated_fwd(feat, rois, (7, 7), 0.25, 2, version)
        _check(got, want, 1e-5, f"pixel-major v{version}")
        assert np.abs(got[5]).max() == 0.0


def test_fused_extractor_full_size_subset(cuda, oracle):
    """BASELINE config 2 / config 1 sizes (1024^2 tile, C=256, K=4000 / K=2000): the real launch, checked against
    the oracle on a random 200-RoI subset (RoIs are independent, so the oracle runs on the subset only)."""
    from rs_detection_b200 import core
    import bench as B
    fs, rois, _, _ = B.tile_inputs(0)
    shapes = [f.shape for f in fs]
    cfg = core.make_roi_cfg(shapes, [1 / s for s in W.STRIDES], 7, 2, 1, B.EXTEND, 56.0)
    xs = [_t(f) for f in fs]
    for K in (4000, 2000):
        r = rois[:K]
        got = core.roi_align_rotated_forward(cfg, xs, _t(r))
        sub = np.sort(np.random.default_rng(K).choice(K, 200, replace=False))
        want, _ = oracle.oriented_extractor_fwd(fs, r[sub], list(W.STRIDES), extend_factor=B.EXTEND)
        _check(got[torch.from_numpy(sub).cuda()].cpu().numpy(), want, 1e-5, f"full-size extractor K={K} (200-RoI subset)")


def test_persistent_kernel_two_chunks_many_items(cuda, oracle):
    """The persistent 7x7 kernel with C = 512 (two 256-channel items per RoI) and more items than resident CTAs
    (148 x 3): every CTA loops, items of one RoI may land on different CTAs.  NCHW and channels-last callers."""
    from rs_detection_b200 import core
    tile, C, K = 128, 512, 700
    shapes = W.fpn_shapes(1, tile=tile, channels=C)
    rng = np.random.default_rng(5)
    feats = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    rois = W.proposals(K, 21, canvas=tile)
    want, _ = oracle.oriented_extractor_fwd(feats, rois, [4, 8, 16, 32])
    for cl in (False, True):
        cfg = core.make_roi_cfg(shapes, [1 / s for s in W.STRIDES], 7, 2, 1, (1.4, 1.2), 56.0, channels_last=cl)
        f = [_t(x) for x in feats]
        if cl:
            f = [core.nchw_to_nhwc(x) for x in f]
        got = core.roi_align_rotated_forward(cfg, f, _t(rois)).cpu().numpy()
        _check(got, want, 1e-5, f"persistent kernel, C=512, channels_last={cl}")
