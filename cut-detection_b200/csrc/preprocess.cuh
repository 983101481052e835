// Resize plan shared between the preprocessing kernels and the fused network entry point.
#pragma once

#include <vector>

#include "common.cuh"

namespace cutdet {

enum { RESIZE_LINEAR = 0, RESIZE_AREA2 = 1, RESIZE_COPY = 2 };

// Passed to kernels by value: geometry + device pointers to the tap tables.
struct ResizePlanDev {
    int src_h, src_w, dst_h, dst_w;
    int mode;
    int all_a1_zero, all_b1_zero;
    // non-zero when the resize degenerates to out[y][x] = src[off_y + y*step_y][off_x + x*step_x]
    int gather_step_x, gather_step_y, gather_off_x, gather_off_y;
    const int *x0, *x1, *a0, *a1;   // [dst_w]
    const int *y0, *y1, *b0, *b1;   // [dst_h]
    const int *row_slot;            // [src_h] source row -> slot in a row-compacted frame
    const int2 *xpack;              // [dst_w] two-tap resizes: {3 * x0, a0 | a1 << 16}, a clamped edge folded into a0 (16-byte aligned)
};

int check_frames(const struct ::cutdet_resize_plan *plan, const cutdet_frames *src);

}  // namespace cutdet

struct cutdet_resize_plan {
    cutdet::ResizePlanDev host;
    void *dev_blob = nullptr;
    std::vector<int> rows;   // source rows the resize reads
    int n_rows = 0;
    // every output row reads source row y0 and (with a non-zero weight) only the row right after it, which is also the next
    // compact slot: the two rows can then be fetched as ONE two-row box (conv12_frames' TMA loaders)
    bool pair_rows = false;
};
