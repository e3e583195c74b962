"""jdet.models.roi_heads.oriented_rpn_head -- only the PROPOSAL STAGE of OrientedRPNHead:
`_get_bboxes_single` / `get_bboxes` (python/jdet/models/roi_heads/oriented_rpn_head.py:136-216, 218-245) as one
device pipeline (`rsdet_rpn_proposals`).  The convolutions, losses and training targets stay in Jittor."""
from .... import core
from ...ops._io import back, dev
from ..boxes.anchor_generator import AnchorGenerator


class OrientedRPNProposals:
    """Carries what `_get_bboxes_single` reads from the head (`nms_pre`, `nms_post`, `nms_thresh`,
    `min_bbox_size`, `use_sigmoid_cls`, the MidpointOffsetCoder means/stds and the anchor generator;
    oriented_rpn_head.py:20-45, configs/orcnn_van3_for_test_1.py:20-45)."""

    def __init__(self, min_bbox_size=0, nms_thresh=0.8, nms_pre=2000, nms_post=2000, use_sigmoid_cls=True,
                 anchor_generator=None, target_means=(.0, .0, .0, .0, .0, .0), target_stds=(1.0, 1.0, 1.0, 1.0, 0.5, 0.5)):
        self.min_bbox_size, self.nms_thresh, self.nms_pre, self.nms_post = min_bbox_size, nms_thresh, nms_pre, nms_post
        self.use_sigmoid_cls = use_sigmoid_cls
        self.means, self.stds = tuple(target_means), tuple(target_stds)
        self.anchor_generator = anchor_generator or AnchorGenerator(strides=[4, 8, 16, 32, 64], ratios=[0.5, 1.0, 2.0], scales=[8])
        self.num_anchors = self.anchor_generator.num_base_anchors[0]

    def _get_bboxes_single(self, cls_scores, bbox_preds, mlvl_anchors, img_shape=None):
        """cls_scores[l] (A*c, H, W), bbox_preds[l] (A*6, H, W), mlvl_anchors[l] (H*W*A, 4) -> dets (k,6)
        [cx,cy,w,h,theta,score], k <= nms_post, descending score.  `img_shape` is ignored like in the reference
        (MidpointOffsetCoder.decode never clips, coder.py:383-433)."""
        cs, fl = zip(*[dev(t) for t in cls_scores])
        bp = [dev(t)[0] for t in bbox_preds]
        an = [dev(t)[0] for t in mlvl_anchors]
        dets, cnt = core.rpn_proposals(cs, bp, an, self.num_anchors, self.use_sigmoid_cls, self.nms_pre, self.nms_post,
                                       self.nms_thresh, self.min_bbox_size, self.means, self.stds)
        return back(dets[:int(cnt.item())], fl[0])

    def get_bboxes(self, cls_scores, bbox_preds, targets=None):
        """cls_scores[l] (N, A*c, H, W) -> list of per-image dets (:218-245)."""
        num_levels = len(cls_scores)
        sizes = [tuple(cls_scores[i].shape[-2:]) for i in range(num_levels)]
        anchors = self.anchor_generator.grid_anchors(sizes)
        out = []
        for img_id in range(cls_scores[0].shape[0]):
            out.append(self._get_bboxes_single([cls_scores[i][img_id] for i in range(num_levels)],
                                               [bbox_preds[i][img_id] for i in range(num_levels)], anchors))
        return out
