"""GPU parity: box_iou_rotated(_v1) + MaxIoU assignment through the jdet mirror -> C ABI vs the oracle.
Contract (BASELINE.json north_star): assignment labels bit-exact outside the 1e-6 IoU band."""
import numpy as np
import pytest
import torch

import workloads as W
from helpers import band_pairs

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("version", [0, 1])
@pytest.mark.parametrize("G", [8, 64, 512])
def test_iou_matrix_vs_oracle(cuda, oracle, version, G):
    from rs_detection_b200.jdet.ops import box_iou_rotated, box_iou_rotated_v1
    P = W.rotated_boxes(2000, 10 + G)
    gt = W.jittered_copies(P, G, 11 + G)
    fn = box_iou_rotated_v1 if version else box_iou_rotated
    got = fn(_t(gt), _t(P)).cpu().numpy()
    want = oracle.box_iou_rotated(gt, P, version, 1)
    assert got.shape == (G, 2000)
    nbad = int((got != want).sum())
    print(f"iou v{version} G={G}: {nbad} of {got.size} values not bit-identical, max|d|={np.abs(got - want).max():.3g}")
    assert np.abs(got - want).max() <= 1e-6
    assert nbad <= got.size * 1e-4


def test_iou_known_answer_and_host_inputs(cuda):
    from rs_detection_b200.jdet.ops import box_iou_rotated
    b = np.array([[0, 0, 1, 1, 0], [0.5, 0.5, 1, 2, 0]], np.float32)
    out = box_iou_rotated(b, b)  # numpy in -> numpy out (host buffers, copies inside)
    assert isinstance(out, np.ndarray)
    np.testing.assert_allclose(out, [[1, 0.2], [0.2, 1]], atol=1e-7)


def test_iou_edge_cases(cuda, oracle):
    from rs_detection_b200.jdet.ops import box_iou_rotated, box_iou_rotated_v1
    b = np.array([[5, 5, 4, 2, 0.3], [5, 5, 4, 2, 0.3], [5, 5, 0, 2, 0], [5, 5, 1e-8, 1e-8, 0], [9, 5, 4, 2, 0.3],
                  [5, 5, 2, 4, 0.3 + np.pi / 2], [7, 5, 4, 2, 0.3], [5, 5, 4, 2, -0.3], [5, 5, 0.0005, 3, 0.1]], np.float32)
    assert np.array_equal(box_iou_rotated(_t(b), _t(b)).cpu().numpy(), oracle.box_iou_rotated(b, b, 0, 1))
    assert np.array_equal(box_iou_rotated_v1(_t(b), _t(b)).cpu().numpy(), oracle.box_iou_rotated_v1(b, b, 1))
    # empty inputs
    e = torch.zeros((0, 5), device="cuda")
    assert box_iou_rotated(e, _t(b)).shape == (0, 9)
    assert box_iou_rotated(_t(b), e).shape == (9, 0)


@pytest.mark.parametrize("G", [8, 64, 512])
def test_assignment_labels_bit_exact(cuda, oracle, G):
    """config 3: box_iou_rotated_v1(gt, proposals) + MaxIoUAssigner(0.5/0.5, match_low_quality=False)."""
    from rs_detection_b200.jdet.models.boxes.assigner import MaxIoUAssigner
    P = W.rotated_boxes(2000, 20 + G)
    gt = W.jittered_copies(P, G, 21 + G)
    gl = np.random.default_rng(G).integers(0, 15, G).astype(np.int32)
    want_ov = oracle.box_iou_rotated_v1(gt, P, 1)
    w_inds, w_max, w_lab = oracle.max_iou_assign(want_ov, 0.5, 0.5, 0.5, False, gt_labels=gl)
    a = MaxIoUAssigner(0.5, 0.5, 0.5, match_low_quality=False, ignore_iof_thr=-1,
                       iou_calculator=dict(type='BboxOverlaps2D_rotated_v1'))
    res = a.assign(_t(P), _t(gt), gt_labels=_t(gl))
    nband = band_pairs(want_ov, 0.5)
    print(f"G={G}: pairs inside the 1e-6 band around 0.5: {nband}")
    got = res.gt_inds.cpu().numpy()
    if nband == 0:
        assert np.array_equal(got, w_inds)
        assert np.array_equal(res.labels.cpu().numpy(), w_lab)
    else:
        cols = np.nonzero((np.abs(want_ov.astype(np.float64) - 0.5) <= 1e-6).any(0))[0]
        ok = np.ones(2000, bool); ok[cols] = False
        assert np.array_equal(got[ok], w_inds[ok])
    assert (got > 0).sum() > 0 and (got == 0).sum() > 0
    np.testing.assert_allclose(res.max_overlaps.cpu().numpy(), w_max, atol=1e-6)


def test_assignment_low_quality_and_range(cuda, oracle):
    from rs_detection_b200 import core
    rng = np.random.default_rng(5)
    ov = rng.uniform(0, 1, (37, 900)).astype(np.float32)
    ov[rng.uniform(size=ov.shape) < 0.7] = 0
    for all_ in (True, False):
        w = oracle.max_iou_assign(ov, 0.7, (0.1, 0.3), 0.3, True, all_)
        g = core.assign_wrt_overlaps(_t(ov), 0.7, (0.1, 0.3), 0.3, True, all_)
        assert np.array_equal(g[0].cpu().numpy(), w[0])
    with pytest.raises(ValueError):
        core.assign_wrt_overlaps(torch.zeros((0, 5), device="cuda"), 0.5, 0.5)
