"""jdet.models.roi_heads.oriented_head -- only the TEST-TIME TAIL of OrientedHead:
`get_bboxes` + `get_results` (python/jdet/models/roi_heads/oriented_head.py:498-536, 279-305) as one fused
device pipeline.  The FC layers, losses and training targets of the head stay in Jittor (out of scope).
"""
from .... import core
from ...ops._io import back, dev


class OrientedHeadTail:
    """Carries the constants `get_bboxes` reads from the head (`bbox_coder` means/stds, `score_thresh`,
    `reg_class_agnostic`, `num_classes`; configs/orcnn_van3_for_test_1.py:66,82-85,108)."""

    def __init__(self, num_classes, score_thresh=0.05, target_means=(0., 0., 0., 0., 0.),
                 target_stds=(0.1, 0.1, 0.2, 0.2, 0.1), reg_class_agnostic=True):
        self.num_classes = num_classes
        self.score_thresh = score_thresh
        self.means, self.stds = tuple(target_means), tuple(target_stds)
        self.reg_class_agnostic = reg_class_agnostic

    def get_bboxes(self, rois, cls_score, bbox_pred, img_shape=None, scale_factor=None, rescale=False):
        """rois (K,6) [batch,cx,cy,w,h,theta], cls_score (K,C+1) logits (background LAST), bbox_pred (K,5) or
        (K,5C) -> (det_bboxes (M,9) [x1..y4,score], det_labels (M,) int64), row-major (roi, class) order.
        `img_shape` is accepted and ignored like in the reference's decode (coder.py:477-514 never clips)."""
        r, fl = dev(rois)
        s, _ = dev(cls_score)
        d, _ = dev(bbox_pred)
        assert s.dim() == 2, "Check cls_score.ndim"
        dets, labels, cnt = core.oriented_head_results(
            r[:, 1:].contiguous(), s, d, self.num_classes, self.reg_class_agnostic, self.means, self.stds,
            self.score_thresh, scale_factor if rescale else None)
        m = int(cnt.item())
        return back(dets[:m], fl), back(labels[:m], fl)
