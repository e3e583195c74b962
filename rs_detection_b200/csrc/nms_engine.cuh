// nms_engine.cuh -- internal interface of the segmented bitmask NMS engine (nms.cu) for the other
// translation units that drive it (rpn.cu).
#pragma once
#include "common.cuh"

namespace rsdet {

struct NmsArgs {
    int kind;
    const void* dets;
    const void* scores;
    const int32_t* labels;
    int n_max;
    const int* n_dev;
    double thr;
    const double* thr_per_label;
    int num_thr;
    uint8_t* keep_mask;
    int64_t* keep_sorted_idx;
    int64_t* keep_score_idx;
    int32_t* num_keep;
    int label_bits = 32;
    // shared-box mode (multiclass with class-agnostic boxes): candidate e refers to box e / cand_per_box
    const float* shared_boxes = nullptr;
    int n_shared = 0;
    int cand_per_box = 1;
    // labels are class ids 0..num_classes-1 whose live counts are already known on the device: the segment
    // table is an exclusive scan of the counts (no boundary search over the sorted labels)
    const int* class_counts = nullptr;
    int num_classes = 0;
    size_t mask_words = 0;  // caller-proved bound on the mask size (0 = worst case n * ceil(n/64))
};

// mask_words = 0: worst case (every box in one label group)
size_t nms_ws_bytes(int kind, int n, size_t mask_words = 0);
int nms_run(const NmsArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace rsdet
