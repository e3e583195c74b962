"""jdet.models.boxes.anchor_generator.AnchorGenerator (python/jdet/models/boxes/anchor_generator.py:93-330),
the subset the oriented RPN uses: `scales` x `ratios` base anchors per stride (scale_major, centre offset 0)
and `grid_anchors`.  Anchors are constants of a config: they are built once on the host and cached on the
device; `rsdet_rpn_proposals` reads them as its `anchors` argument."""
import numpy as np
import torch


class AnchorGenerator:
    def __init__(self, strides, ratios, scales, base_sizes=None):
        self.strides = [int(s) for s in strides]
        self.base_sizes = list(base_sizes) if base_sizes is not None else list(self.strides)
        assert len(self.base_sizes) == len(self.strides)
        self.ratios = np.asarray(ratios, np.float32)
        self.scales = np.asarray(scales, np.float32)
        self.base_anchors = [self._base(b) for b in self.base_sizes]
        self._cache = {}

    @property
    def num_base_anchors(self):
        return [b.shape[0] for b in self.base_anchors]

    @property
    def num_levels(self):
        return len(self.strides)

    def _base(self, base_size):
        h_ratios = np.sqrt(self.ratios)
        w_ratios = (1 / h_ratios).astype(np.float32)
        ws = (np.float32(base_size) * w_ratios[:, None] * self.scales[None, :]).reshape(-1)
        hs = (np.float32(base_size) * h_ratios[:, None] * self.scales[None, :]).reshape(-1)
        return np.stack([-0.5 * ws, -0.5 * hs, 0.5 * ws, 0.5 * hs], -1).astype(np.float32)

    def grid_anchors(self, featmap_sizes, device="cuda"):
        """list of (H*W*A, 4) float32 tensors, first A rows = the A anchors of cell (0,0), then (0,1), ..."""
        assert len(featmap_sizes) == self.num_levels
        out = []
        for l, (h, w) in enumerate(featmap_sizes):
            key = (l, int(h), int(w), str(device))
            if key not in self._cache:
                s = self.strides[l]
                xx = np.tile((np.arange(w) * s).astype(np.float32), h)
                yy = np.repeat((np.arange(h) * s).astype(np.float32), w)
                shifts = np.stack([xx, yy, xx, yy], -1)
                a = (self.base_anchors[l][None] + shifts[:, None]).reshape(-1, 4).astype(np.float32)
                self._cache[key] = torch.from_numpy(a).to(device)
            out.append(self._cache[key])
        return out
