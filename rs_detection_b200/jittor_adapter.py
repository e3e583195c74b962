"""Jittor binding of librsdet.so -- the file a JDet maintainer drops next to `python/jdet/ops/`.

Jittor is NOT installable in this image (no network), so this module cannot be executed here; it is
kept deliberately thin: every op is ONE `jt.code` whose CUDA body resolves the `extern "C"` entry
point of include/rsdet.h with dlopen/dlsym (no extra compile or link flags needed) and forwards
Jittor's raw pointers (`in0_p`, `out0_p`, `in0_shape0`, ... -- the same glue the reference ops use,
e.g. python/jdet/ops/box_iou_rotated.py:464-485) on the legacy default stream, which is where
Jittor launches its own kernels.  Temporary memory comes from Jittor's allocator exactly like the
reference's NMS mask (`exe.allocator->alloc/free`, python/jdet/ops/nms_rotated.py:462-464,492).

Status: UNVERIFIED against a live Jittor ("What still has to be checked on a
Jittor box").  The torch-facing mirror in `rs_detection_b200/jdet/` is what the tests and the
benchmark exercise; both call the same C ABI.
"""
import os

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "librsdet.so")

_HEADER = r'''
#undef out
#include <dlfcn.h>
#include <cstdint>
#include <cstdio>
#include <executor.h>
namespace {
void* rsdet_sym(const char* name) {
  static void* lib = dlopen("%s", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { fprintf(stderr, "librsdet.so: %%s\n", dlerror()); abort(); }
  void* f = dlsym(lib, name);
  if (!f) { fprintf(stderr, "librsdet.so: missing %%s\n", name); abort(); }
  return f;
}
struct Scratch {  // Jittor-owned temporary, as in nms_rotated.py:462-464
  void* p; size_t bytes, alloc;
  explicit Scratch(size_t n) : bytes(n) { p = exe.allocator->alloc(bytes, alloc); }
  ~Scratch() { exe.allocator->free(p, bytes, alloc); }
};
inline void rsdet_check(int rc, const char* what) {
  if (rc != 0) { fprintf(stderr, "%%s failed: rsdet code %%d\n", what, rc); abort(); }
}
}
''' % LIB


def _jt():
    import jittor as jt  # noqa: raises ModuleNotFoundError where Jittor is absent (this image)
    return jt


def box_iou_rotated(boxes1, boxes2, version=0):
    """jdet.ops.box_iou_rotated / box_iou_rotated_v1 (box_iou_rotated.py:502-509, box_iou_rotated_v1.py:507-524)."""
    jt = _jt()
    assert boxes1.dtype == boxes2.dtype
    n1, n2 = boxes1.shape[0], boxes2.shape[0]
    ious = jt.code((n1 * n2,), boxes1.dtype, [boxes1, boxes2], cuda_header=_HEADER, cuda_src=r'''
        typedef size_t (*ws_t)(int, int);
        typedef int (*fn_t)(const float*, int, const float*, int, int, int, float*, void*, size_t, void*);
        static ws_t ws = (ws_t)rsdet_sym("rsdet_box_iou_rotated_workspace_bytes");
        static fn_t fn = (fn_t)rsdet_sym("rsdet_box_iou_rotated");
        int n1 = in0_shape0, n2 = in1_shape0;
        if (n1 > 0 && n2 > 0) {
          Scratch s(ws(n1, n2));
          rsdet_check(fn(in0_p, n1, in1_p, n2, %d, %d, out0_p, s.p, s.bytes, 0), "box_iou_rotated");
        }''' % (version, 1 if version == 1 else 0))
    return ious.reshape(n1, n2)


def box_iou_rotated_v1(boxes1, boxes2):
    return box_iou_rotated(boxes1, boxes2, 1)


def _nms_keep(kind, dets, scores, labels, thr):
    """bool keep mask in ORIGINAL index space (nms_rotated_cuda's `keep`, nms_rotated.py:506-513)."""
    jt = _jt()
    ins = [dets, scores] + ([labels.int32()] if labels is not None else [])
    return jt.code((dets.shape[0],), "uint8", ins, cuda_header=_HEADER, cuda_src=r'''
        typedef size_t (*ws_t)(int, int);
        typedef int (*fn_t)(int, const void*, const void*, const int32_t*, int, double, const double*, int, uint8_t*,
                            int64_t*, int64_t*, int32_t*, void*, size_t, void*);
        static ws_t ws = (ws_t)rsdet_sym("rsdet_nms_workspace_bytes");
        static fn_t fn = (fn_t)rsdet_sym("rsdet_nms");
        int n = in0_shape0;
        cudaMemsetAsync(out0_p, 0, out0->size);
        if (n > 0) {
          Scratch s(ws(%d, n));
          rsdet_check(fn(%d, in0_p, in1_p, %s, n, %r, nullptr, 0, (uint8_t*)out0_p, nullptr, nullptr, nullptr,
                         s.p, s.bytes, 0), "nms");
        }''' % (kind, kind, "(const int32_t*)in2_p" if labels is not None else "nullptr", float(thr))).bool()


def nms_rotated(dets, scores, iou_threshold):
    """nms_rotated.py:527-538"""
    jt = _jt()
    if dets.numel() == 0:
        return jt.array([])
    assert dets.numel() > 0 and dets.ndim == 2 and dets.dtype == scores.dtype
    return jt.where(_nms_keep(0, dets, scores, None, iou_threshold))[0]


def ml_nms_rotated(dets, scores, labels, iou_threshold):
    """nms_rotated.py:515-525"""
    jt = _jt()
    assert dets.numel() > 0 and dets.ndim == 2 and dets.dtype == scores.dtype
    return jt.where(_nms_keep(0, dets, scores, labels, iou_threshold))[0]


def poly_nms(boxes, nms_overlap_thresh):
    """nms_poly.py:187-232 -> order_t[keep]"""
    jt = _jt()
    assert boxes.ndim == 2 and boxes.shape[1] == 9
    scores = boxes[:, 8]
    order_t, _ = scores.argsort(0, descending=True)
    keep = _nms_keep(2, boxes[:, :8], scores, None, nms_overlap_thresh)
    return order_t[keep[order_t]]


_ROI_CFG = r'''
        struct Cfg { int num_levels, batch, channels; int height[8], width[8]; float spatial_scale[8];
                     int pooled_h, pooled_w, sampling_ratio, version; float extend_w, extend_h, finest_scale;
                     int channels_last; };
        Cfg cfg; memset(&cfg, 0, sizeof cfg);
        cfg.num_levels = 1; cfg.batch = in0_shape0; cfg.channels = in0_shape1;
        cfg.height[0] = in0_shape2; cfg.width[0] = in0_shape3; cfg.spatial_scale[0] = %r;
        cfg.pooled_h = %d; cfg.pooled_w = %d; cfg.sampling_ratio = %d; cfg.version = %d;
        cfg.extend_w = 1.f; cfg.extend_h = 1.f; cfg.finest_scale = 56.f;
        typedef size_t (*ws_t)(const Cfg*, int, int);
        static ws_t ws = (ws_t)rsdet_sym("rsdet_roi_align_rotated_workspace_bytes");
'''


def make_roi_align(version):
    """Returns the jt.Function class replacing _RotatedROIAlign(_v1) (roi_align_rotated_v1.py:300-353)."""
    jt = _jt()

    class _Fn(jt.Function):
        def execute(self, input, rois, output_size, spatial_scale, sampling_ratio):
            self.input, self.rois = input, rois
            self.args = (float(spatial_scale), int(output_size[0]), int(output_size[1]), int(sampling_ratio), version)
            assert rois.shape[1] == 6
            shape = (rois.shape[0], input.shape[1], output_size[0], output_size[1])
            return jt.code(shape, input.dtype, [input, rois], cuda_header=_HEADER, cuda_src=(_ROI_CFG % self.args) + r'''
        typedef int (*fn_t)(const Cfg*, const float* const*, const float*, int, float*, int32_t*, void*, size_t, void*);
        static fn_t fn = (fn_t)rsdet_sym("rsdet_roi_align_rotated_forward");
        int K = in1_shape0;
        if (K > 0) {
          Scratch s(ws(&cfg, K, 0));
          const float* feats[8] = {in0_p};
          rsdet_check(fn(&cfg, feats, in1_p, K, out0_p, nullptr, s.p, s.bytes, 0), "roi_align_rotated_forward");
        }''')

        def grad(self, output_grad):
            input, rois = self.input, self.rois
            g = jt.code(input.shape, input.dtype, [input, rois, output_grad], cuda_header=_HEADER,
                        cuda_src=(_ROI_CFG % self.args) + r'''
        typedef int (*fn_t)(const Cfg*, const float*, const float*, int, float* const*, void*, size_t, void*);
        static fn_t fn = (fn_t)rsdet_sym("rsdet_roi_align_rotated_backward");
        int K = in1_shape0;
        Scratch s(ws(&cfg, K, 1));
        float* grads[8] = {out0_p};
        rsdet_check(fn(&cfg, in2_p, in1_p, K, grads, s.p, s.bytes, 0), "roi_align_rotated_backward");''')
            return g, None

    return _Fn


def multiclass_nms_rotated(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None):
    """jdet.ops.nms_rotated.multiclass_nms_rotated (nms_rotated.py:540-596) as ONE jt.code: candidate selection,
    per-class NMS through the shared decision matrix, score sort and top-k on the device.  Returns
    (dets (k,6), labels (k,)) like the reference; k is read back once (the reference syncs in jt.where)."""
    jt = _jt()
    n, C = multi_scores.shape[0], multi_scores.shape[1] - 1
    if n == 0:
        return jt.zeros((0, 6)), jt.zeros((0,)).int32()
    cap = n * C  # the C entry point needs n*C rows of output whatever max_num is
    ins = [multi_bboxes, multi_scores] + ([score_factors] if score_factors is not None else [])
    dets, labels, cnt = jt.code([(cap, 6), (cap,), (1,)], [multi_bboxes.dtype, "int32", "int32"], ins, cuda_header=_HEADER,
                                cuda_src=r'''
        typedef size_t (*ws_t)(int, int);
        typedef int (*fn_t)(const float*, int, const float*, int, int, float, float, int, const float*, float*, int32_t*,
                            int32_t*, void*, size_t, void*);
        static ws_t ws = (ws_t)rsdet_sym("rsdet_multiclass_nms_rotated_workspace_bytes");
        static fn_t fn = (fn_t)rsdet_sym("rsdet_multiclass_nms_rotated");
        int n = in1_shape0, C = in1_shape1 - 1;
        Scratch s(ws(n, C));
        rsdet_check(fn(in0_p, in0_shape1, in1_p, n, C, %r, %r, %d, %s, out0_p, (int32_t*)out1_p, (int32_t*)out2_p,
                       s.p, s.bytes, 0), "multiclass_nms_rotated");''' % (float(score_thr), float(nms_cfg.get('iou_thr', 0.1)),
                                                                          int(max_num), "in2_p" if score_factors is not None else "nullptr"))
    k = int(cnt.item())
    return dets[:k], labels[:k]


def rpn_proposals(cls_scores, bbox_preds, mlvl_anchors, num_anchors, use_sigmoid=True, nms_pre=2000, nms_post=2000,
                  nms_thresh=0.8, min_bbox_size=0, means=(0.,) * 6, stds=(1., 1., 1., 1., 0.5, 0.5)):
    """OrientedRPNHead._get_bboxes_single (oriented_rpn_head.py:136-216) as one jt.code over 3L inputs."""
    jt = _jt()
    L = len(cls_scores)
    ins = list(cls_scores) + list(bbox_preds) + list(mlvl_anchors)
    lv = "".join("cfg.height[%d] = in%d_shape1; cfg.width[%d] = in%d_shape2; cls[%d] = in%d_p; reg[%d] = in%d_p; anc[%d] = in%d_p;\n"
                 % (l, l, l, l, l, l, l, L + l, l, 2 * L + l) for l in range(L))
    dets, cnt = jt.code([(nms_post, 6), (1,)], [cls_scores[0].dtype, "int32"], ins, cuda_header=_HEADER, cuda_src=r'''
        struct Cfg { int num_levels; int height[8], width[8]; int num_anchors, use_sigmoid, nms_pre, nms_post; double nms_thresh;
                     float min_bbox_size, means[6], stds[6], wh_ratio_clip; };
        Cfg cfg; memset(&cfg, 0, sizeof cfg);
        const float *cls[8], *reg[8], *anc[8];
        cfg.num_levels = %d; cfg.num_anchors = %d; cfg.use_sigmoid = %d; cfg.nms_pre = %d; cfg.nms_post = %d;
        cfg.nms_thresh = %r; cfg.min_bbox_size = %r; cfg.wh_ratio_clip = 0.016f;
        const float mm[6] = {%s}, ss[6] = {%s};
        for (int k = 0; k < 6; k++) { cfg.means[k] = mm[k]; cfg.stds[k] = ss[k]; }
        %s
        typedef size_t (*ws_t)(const Cfg*);
        typedef int (*fn_t)(const Cfg*, const float* const*, const float* const*, const float* const*, float*, int32_t*, float*,
                            float*, float*, int32_t*, void*, size_t, void*);
        static ws_t ws = (ws_t)rsdet_sym("rsdet_rpn_proposals_workspace_bytes");
        static fn_t fn = (fn_t)rsdet_sym("rsdet_rpn_proposals");
        Scratch s(ws(&cfg));
        rsdet_check(fn(&cfg, cls, reg, anc, out0_p, (int32_t*)out1_p, nullptr, nullptr, nullptr, nullptr, s.p, s.bytes, 0),
                    "rpn_proposals");''' % (L, num_anchors, int(use_sigmoid), nms_pre, nms_post, float(nms_thresh), float(min_bbox_size),
                                           ", ".join("%rf" % float(m) for m in means), ", ".join("%rf" % float(v) for v in stds), lv))
    return dets[:int(cnt.item())]
