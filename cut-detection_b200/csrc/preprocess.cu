// K1: decoded uint8 BGR HWC frames -> resized / normalised model input.
//
// Stands in for VideoDataset.__next__ (reference frameID/data.py:218-228):
//     cv2.resize(frame, (W2, H2), INTER_LINEAR) -> float -> permute(2,0,1) -> flip(BGR->RGB) -> /255
// The resize arithmetic is OpenCV's 11-bit fixed point for uint8 images (see oracle/preprocess.py for the
// restatement it is tested against, bit for bit):
//     S   = a0 * p[x0] + a1 * p[x0 + 1]                                  (horizontal, int32)
//     out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2   (vertical)
// with an exact 2x2 box mean when both scales are exactly 2, and a plain copy at scale 1.
//
// The tap tables are built on the host with the same float32 steps OpenCV takes and live in a
// cutdet_resize_plan.  Taps whose weight is zero are never loaded: at 1280x720 -> 256x144 (scale 5) every
// second tap has weight 0, so the kernel touches only source rows 5y+2 -- 144 of the 720 rows.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "preprocess.cuh"

namespace cutdet {

namespace {

void linear_taps(int src, int dst, bool clamp_weights, std::vector<int> &i0, std::vector<int> &i1,
                 std::vector<int> &w0, std::vector<int> &w1) {
    i0.resize(dst); i1.resize(dst); w0.resize(dst); w1.resize(dst);
    const double scale = 1.0 / ((double)dst / (double)src);
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (clamp_weights) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= src - 1) { f = 0.f; s = src - 1; }
        }
        w0[d] = (int)lrintf((1.f - f) * 2048.f);   // round half to even, like cvRound
        w1[d] = (int)lrintf(f * 2048.f);
        int a = s < 0 ? 0 : (s > src - 1 ? src - 1 : s);
        int b = s + 1 < 0 ? 0 : (s + 1 > src - 1 ? src - 1 : s + 1);
        i0[d] = a;
        i1[d] = b;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------- kernels
// One thread per output pixel, x fastest.  OUT: 0 = float32 NCHW RGB /255, 1 = uint8 HWC BGR.
template <int OUT>
__global__ void __launch_bounds__(256) preprocess_generic_kernel(ResizePlanDev plan, const uint8_t *__restrict__ frames,
                                                                 int64_t frame_stride, int64_t row_pitch, int compact,
                                                                 void *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= plan.dst_w || y >= plan.dst_h) return;
    const uint8_t *frame = frames + (int64_t)b * frame_stride;
    int v[3];
    if (plan.gather_step_x > 0) {
        // integer scale (720p, 1440p, 2160p -> 256 wide): every second tap has zero weight, out[y][x] = src[off_y + y sy][off_x + x sx].
        // No tap tables; the three bytes come from two aligned 32-bit loads (rows and frames are 4-byte aligned or we fall back).
        const int sy = plan.gather_off_y + y * plan.gather_step_y;
        const int r = compact ? plan.row_slot[sy] : sy;
        const uint8_t *p = frame + (int64_t)r * row_pitch + 3 * (plan.gather_off_x + x * plan.gather_step_x);
        if (((frame_stride | row_pitch | reinterpret_cast<uintptr_t>(frames)) & 3) == 0) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(p);
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
            const uint32_t lo = __ldg(wp), hi = (a & 3) > 1 ? __ldg(wp + 1) : 0u;     // the 2nd word only when the pixel straddles it
            const uint32_t px = __funnelshift_r(lo, hi, (uint32_t)(a & 3) * 8);
            v[0] = px & 0xff; v[1] = (px >> 8) & 0xff; v[2] = (px >> 16) & 0xff;
        } else {
            v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
        }
    } else if (plan.mode == RESIZE_COPY) {
        const int r = compact ? plan.row_slot[y] : y;
        const uint8_t *p = frame + (int64_t)r * row_pitch + 3 * x;
        v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
    } else if (plan.mode == RESIZE_AREA2) {
        const int r0 = compact ? plan.row_slot[2 * y] : 2 * y;
        const int r1 = compact ? plan.row_slot[2 * y + 1] : 2 * y + 1;
        const uint8_t *p0 = frame + (int64_t)r0 * row_pitch + 6 * x;
        const uint8_t *p1 = frame + (int64_t)r1 * row_pitch + 6 * x;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2;
    } else {
        const int x0 = plan.x0[x], x1 = plan.x1[x], a0 = plan.a0[x], a1 = plan.a1[x];
        const int y0 = plan.y0[y], y1 = plan.y1[y], b0 = plan.b0[y], b1 = plan.b1[y];
        const int r0 = compact ? plan.row_slot[y0] : y0;
        const uint8_t *p0 = frame + (int64_t)r0 * row_pitch;
        int s0[3], s1[3] = {0, 0, 0};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int s = a0 * p0[3 * x0 + c];
            if (a1 != 0) s += a1 * p0[3 * x1 + c];
            s0[c] = s;
        }
        if (b1 != 0) {
            const int r1 = compact ? plan.row_slot[y1] : y1;
            const uint8_t *p1 = frame + (int64_t)r1 * row_pitch;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int s = a0 * p1[3 * x0 + c];
                if (a1 != 0) s += a1 * p1[3 * x1 + c];
                s1[c] = s;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int o = (((b0 * (s0[c] >> 4)) >> 16) + ((b1 * (s1[c] >> 4)) >> 16) + 2) >> 2;
            v[c] = min(max(o, 0), 255);
        }
    }
    if (OUT == 0) {
        float *o = reinterpret_cast<float *>(out);
        const int64_t plane = (int64_t)plan.dst_h * plan.dst_w;
        const int64_t base = (int64_t)b * 3 * plane + (int64_t)y * plan.dst_w + x;
        // BGR -> RGB: output channel 0 is source channel 2.  True float32 division, as the reference does.
        o[base] = __fdiv_rn((float)v[2], 255.f);
        o[base + plane] = __fdiv_rn((float)v[1], 255.f);
        o[base + 2 * plane] = __fdiv_rn((float)v[0], 255.f);
    } else {
        uint8_t *o = reinterpret_cast<uint8_t *>(out) + ((int64_t)b * plan.dst_h * plan.dst_w + (int64_t)y * plan.dst_w + x) * 3;
        o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
    }
}

// The same arithmetic, one block per (frame, output row), for 16-byte aligned frames (every common resolution: 3 * width is a
// multiple of 16).  The one or two source rows the output row reads come in with 128-bit loads (coalesced, read-only path,
// all of a block's loads in flight at once), the taps are taken from shared memory (two funnel-shifted words per pixel for a
// gather, DP2A on byte pairs for the bilinear case), and the output row leaves through shared memory as 128-bit stores --
// float4 per channel plane, or uint4 over the packed BGR bytes.
constexpr int ROWS_THREADS = 128;
constexpr int ROWS_PER_BLOCK = 2;       // output rows per block: all their source rows are in flight together (cp.async)

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One resized pixel (B, G, R, 0) of output row `y`, column x, from the staged source rows.
__device__ __forceinline__ uint32_t rows_pixel(const ResizePlanDev &plan, const uint8_t *s_row0, const uint8_t *s_row1, bool two, int b0,
                                               int b1, int x) {
    if (plan.gather_step_x > 0 || plan.mode == RESIZE_COPY) {
        const int o = plan.gather_step_x > 0 ? 3 * (plan.gather_off_x + x * plan.gather_step_x) : 3 * x;
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(s_row0 + (o & ~3));
        return __funnelshift_r(wp[0], wp[1], (o & 3) * 8) & 0x00FFFFFFu;
    }
    if (plan.mode == RESIZE_AREA2) {
        const int o = 6 * x;
        const uint32_t *w0 = reinterpret_cast<const uint32_t *>(s_row0 + (o & ~3));
        const uint32_t *w1 = reinterpret_cast<const uint32_t *>(s_row1 + (o & ~3));
        const uint32_t sh = (o & 3) * 8;
        const uint32_t lo0 = __funnelshift_r(w0[0], w0[1], sh), hi0 = __funnelshift_r(w0[1], w0[2], sh);   // bytes 0..3, 4..7
        const uint32_t lo1 = __funnelshift_r(w1[0], w1[1], sh), hi1 = __funnelshift_r(w1[1], w1[2], sh);
        uint32_t px = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t selc = c == 0 ? 0x7730u : (c == 1 ? 0x7741u : 0x7752u);       // (p[2x][c], p[2x+1][c]) in the low half
            const uint32_t v = (__dp2a_lo(0x00010001u, __byte_perm(lo0, hi0, selc), 0u) +
                                __dp2a_lo(0x00010001u, __byte_perm(lo1, hi1, selc), 0u) + 2u) >> 2;
            px |= v << (8 * c);
        }
        return px;
    }
    const int x0 = __ldg(plan.x0 + x), x1 = __ldg(plan.x1 + x);
    int a0 = __ldg(plan.a0 + x), a1 = __ldg(plan.a1 + x);
    if (x1 != x0 + 1) { a0 += a1; a1 = 0; }            // clamped at the edge: both taps are the same pixel
    const uint32_t aw = (uint32_t)a0 | ((uint32_t)a1 << 16);
    const int o = 3 * x0;
    const uint32_t sh = (o & 3) * 8;
    const uint32_t *w0 = reinterpret_cast<const uint32_t *>(s_row0 + (o & ~3));
    const uint32_t lo0 = __funnelshift_r(w0[0], w0[1], sh), hi0 = __funnelshift_r(w0[1], w0[2], sh);
    uint32_t lo1 = 0, hi1 = 0;
    if (two) {
        const uint32_t *w1 = reinterpret_cast<const uint32_t *>(s_row1 + (o & ~3));
        lo1 = __funnelshift_r(w1[0], w1[1], sh); hi1 = __funnelshift_r(w1[1], w1[2], sh);
    }
    uint32_t px = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t selc = c == 0 ? 0x7730u : (c == 1 ? 0x7741u : 0x7752u);
        const int s0 = (int)__dp2a_lo(aw, __byte_perm(lo0, hi0, selc), 0u);
        const int s1 = two ? (int)__dp2a_lo(aw, __byte_perm(lo1, hi1, selc), 0u) : 0;
        const int o8 = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
        px |= (uint32_t)min(max(o8, 0), 255) << (8 * c);
    }
    return px;
}

template <int OUT>
__global__ void __launch_bounds__(ROWS_THREADS) preprocess_rows_kernel(ResizePlanDev plan, const uint8_t *__restrict__ frames,
                                                                      int64_t frame_stride, int64_t row_pitch, int compact,
                                                                      void *__restrict__ out, int row_pad, int out_pad, int vec_out) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int y0 = blockIdx.x * ROWS_PER_BLOCK, b = blockIdx.y, tid = threadIdx.x;
    const int row_bytes = 3 * plan.src_w, n16 = (row_bytes + 15) >> 4;
    const uint8_t *frame = frames + (int64_t)b * frame_stride;
    uint8_t *s_out = smem + 2 * ROWS_PER_BLOCK * row_pad;
    int yb0[ROWS_PER_BLOCK], yb1[ROWS_PER_BLOCK];
    // stage: every source row of the block's output rows, 16 bytes per cp.async, all in flight before the one wait
#pragma unroll
    for (int r = 0; r < ROWS_PER_BLOCK; ++r) {
        const int y = y0 + r;
        yb0[r] = 2048; yb1[r] = 0;
        if (y >= plan.dst_h) continue;
        int r0, r1;
        if (plan.gather_step_x > 0) { r0 = r1 = plan.gather_off_y + y * plan.gather_step_y; }
        else if (plan.mode == RESIZE_COPY) { r0 = r1 = y; }
        else if (plan.mode == RESIZE_AREA2) { r0 = 2 * y; r1 = 2 * y + 1; yb1[r] = 1; }
        else { r0 = plan.y0[y]; r1 = plan.y1[y]; yb0[r] = plan.b0[y]; yb1[r] = plan.b1[y]; }
        const bool two = yb1[r] != 0;
        if (compact) { r0 = plan.row_slot[r0]; r1 = two ? plan.row_slot[r1] : r0; }
        const uint4 *g0 = reinterpret_cast<const uint4 *>(frame + (int64_t)r0 * row_pitch);
        const uint4 *g1 = reinterpret_cast<const uint4 *>(frame + (int64_t)r1 * row_pitch);
        uint4 *d0 = reinterpret_cast<uint4 *>(smem + (2 * r) * row_pad), *d1 = reinterpret_cast<uint4 *>(smem + (2 * r + 1) * row_pad);
        for (int i = tid; i < n16; i += ROWS_THREADS) {
            cp_async_16(d0 + i, g0 + i);
            if (two) cp_async_16(d1 + i, g1 + i);
        }
        if (tid == 0) {        // the taps of the last pixels read a word or two past the row
            d0[n16] = make_uint4(0, 0, 0, 0);
            d1[n16] = make_uint4(0, 0, 0, 0);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    const int dst_w = plan.dst_w;
#pragma unroll
    for (int r = 0; r < ROWS_PER_BLOCK; ++r) {
        if (y0 + r >= plan.dst_h) break;
        const uint8_t *s_row0 = smem + (2 * r) * row_pad, *s_row1 = smem + (2 * r + 1) * row_pad;
        uint8_t *so8 = s_out + r * out_pad;
        for (int x = tid; x < dst_w; x += ROWS_THREADS) {
            const uint32_t px = rows_pixel(plan, s_row0, s_row1, yb1[r] != 0, yb0[r], yb1[r], x);
            if (OUT == 0) {
                float *so = reinterpret_cast<float *>(so8);      // [3][dst_w], RGB: output channel 0 is source channel 2
                so[x] = __fdiv_rn((float)((px >> 16) & 0xff), 255.f);
                so[dst_w + x] = __fdiv_rn((float)((px >> 8) & 0xff), 255.f);
                so[2 * dst_w + x] = __fdiv_rn((float)(px & 0xff), 255.f);
            } else {
                so8[3 * x] = (uint8_t)px; so8[3 * x + 1] = (uint8_t)(px >> 8); so8[3 * x + 2] = (uint8_t)(px >> 16);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROWS_PER_BLOCK; ++r) {
        const int y = y0 + r;
        if (y >= plan.dst_h) break;
        const uint8_t *so8 = s_out + r * out_pad;
        if (OUT == 0) {
            const int64_t plane = (int64_t)plan.dst_h * dst_w;
            float *o = reinterpret_cast<float *>(out) + (int64_t)b * 3 * plane + (int64_t)y * dst_w;
            if (vec_out) {
                const int q = dst_w >> 2;
                for (int i = tid; i < 3 * q; i += ROWS_THREADS) {
                    const int c = i / q, j = i - c * q;
                    reinterpret_cast<float4 *>(o + c * plane)[j] = reinterpret_cast<const float4 *>(so8)[c * q + j];
                }
            } else {
                for (int i = tid; i < 3 * dst_w; i += ROWS_THREADS) { const int c = i / dst_w; o[c * plane + (i - c * dst_w)] = reinterpret_cast<const float *>(so8)[i]; }
            }
        } else {
            uint8_t *o = reinterpret_cast<uint8_t *>(out) + ((int64_t)b * plan.dst_h + y) * dst_w * 3;
            if (vec_out) {
                for (int i = tid; i < (3 * dst_w) >> 4; i += ROWS_THREADS) reinterpret_cast<uint4 *>(o)[i] = reinterpret_cast<const uint4 *>(so8)[i];
            } else {
                for (int i = tid; i < 3 * dst_w; i += ROWS_THREADS) o[i] = so8[i];
            }
        }
    }
}

// The resize once more, for all the geometries where instructions -- not the memory -- pace the kernels above (640x360 -> 256x144
// reads 15 source bytes per output pixel where 1080p reads 45; even the 720p gather, which computes nothing, was short of issue
// slots at 78 % of the HBM rate): FOUR ADJACENT output pixels per thread.  The source rows of four output rows are staged as in the row kernel; a thread fetches the packed taps of
// its four columns once (two 16-byte loads of plan.xpack) and uses them for two output rows; the vertical pass needs no clamp
// (the weights of a pair are non-negative and sum to 2048 -- 2049 at most, for a pathological fraction -- and up to that sum in
// both passes the result is at most 255: tests/test_kernel_invariants.py); and the outputs leave straight from
// registers -- a float4 per channel plane (the division by 255 is a 256-entry table of correctly rounded quotients in shared
// memory) or the twelve packed BGR bytes as three words -- with no second pass through shared memory.
// Rows need not be 16-byte aligned (854-pixel rows are 2,562 bytes): a row is staged at its own offset within a 16-byte chunk, as
// whole chunks from the aligned address below it -- up to 15 bytes of its neighbours come along, inside this launch's frames -- and
// only the first and last rows of a batch byte by byte at their ends, so nothing outside the caller's buffer is ever read.
constexpr int QUAD_ROWS = 4;
constexpr int QUAD_THREADS = 128;

// Stage `bytes` bytes from g to slot + (g & 15) (slot 16-byte aligned); zero the 16 bytes behind them (the taps of the last
// pixels read a word or two past the row).  All threads of the block take part.
__device__ __noinline__ void quads_stage_row(uint8_t *slot, const uint8_t *g, int bytes, int tid) {
    const int lead = (int)(reinterpret_cast<uintptr_t>(g) & 15);
    const uint8_t *a = g - lead;                                     // 16-byte aligned
    const int c0 = lead ? 1 : 0, c1 = (lead + bytes) >> 4;           // whole chunks [c0, c1) lie inside the row
    for (int k = c0 + tid; k < c1; k += QUAD_THREADS) cp_async_16(slot + 16 * k, a + 16 * k);
    const int head = lead ? min(16 - lead, bytes) : 0;               // row bytes [0, head) share chunk 0 with what lies before the row
    const int tail0 = max(16 * c1 - lead, head);                     // row bytes [tail0, bytes): the last, partial chunk
    if (tid < head) slot[lead + tid] = __ldg(g + tid);
    const int t = tid - 32;                                          // the second warp: tail bytes, then the zeros behind the row
    if (t >= 0 && t < 32) {
        if (tail0 + t < bytes) slot[lead + tail0 + t] = __ldg(g + tail0 + t);
        if (t < 16) slot[lead + bytes + t] = 0;
    }
}

// MODE: 0 = bilinear taps, 1 = the exact 2x2 mean, 2 = one source pixel per output pixel (integer-scale gathers such as 720p ->
// 256x144, plain copies: one source row per output row, no arithmetic).  ALIGNED: every row starts on a 16-byte boundary (then a
// row sits at the start of its slot and both rows of a pair share the tap offsets).  Compile-time, because the kernel runs within a
// few per cent of the memory roofline only while its instruction count stays where it is: with the 2x2 mean and the alignment as
// run-time switches 640x360 fell from 90 % of the HBM peak to 72 % (profiles/r02_k1_matrix_quads_runtime_switches.txt).
template <int OUT, int MODE, bool ALIGNED>
__global__ void __launch_bounds__(QUAD_THREADS, 10) preprocess_quads_kernel(ResizePlanDev plan, const uint8_t *__restrict__ frames,
                                                                           int64_t frame_stride, int64_t row_pitch, int compact,
                                                                           void *__restrict__ out, int row_pad,
                                                                           const uint8_t *safe_lo, const uint8_t *safe_hi) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int SPR = MODE == 2 ? 1 : 2;          // slots (source rows) per output row
    float *s_lut = reinterpret_cast<float *>(smem + SPR * QUAD_ROWS * row_pad);
    int *s_lead = reinterpret_cast<int *>(smem + SPR * QUAD_ROWS * row_pad + 1024);       // [2 * QUAD_ROWS] where each row starts in its slot
    const int y_first = blockIdx.x * QUAD_ROWS, b = blockIdx.y, tid = threadIdx.x;
    const int row_bytes = 3 * plan.src_w;
    constexpr bool area = MODE == 1, pick = MODE == 2;
    const int n16 = (row_bytes + 15) >> 4;
    const uint8_t *frame = frames + (int64_t)b * frame_stride;
#pragma unroll
    for (int r = 0; r < QUAD_ROWS; ++r) {
        const int y = y_first + r;
        if (y >= plan.dst_h) break;
        int r0, r1;
        bool two;
        if (area) { r0 = 2 * y; r1 = 2 * y + 1; two = true; }
        else if (pick) { r0 = r1 = plan.gather_step_y > 0 ? plan.gather_off_y + y * plan.gather_step_y : y; two = false; }
        else { r0 = __ldg(plan.y0 + y); r1 = __ldg(plan.y1 + y); two = __ldg(plan.b1 + y) != 0; }
        if (compact) { r0 = __ldg(plan.row_slot + r0); r1 = two ? __ldg(plan.row_slot + r1) : r0; }
        const uint8_t *g0 = frame + (int64_t)r0 * row_pitch, *g1 = frame + (int64_t)r1 * row_pitch;
        if (ALIGNED) {
            uint4 *d0 = reinterpret_cast<uint4 *>(smem + (SPR * r) * row_pad), *d1 = reinterpret_cast<uint4 *>(smem + (SPR * r + 1) * row_pad);
            for (int i = tid; i < n16; i += QUAD_THREADS) {
                cp_async_16(d0 + i, reinterpret_cast<const uint4 *>(g0) + i);
                if (!pick && two) cp_async_16(d1 + i, reinterpret_cast<const uint4 *>(g1) + i);
            }
            if (tid == 0) {        // the taps of the last pixels read a word or two past the row
                d0[n16] = make_uint4(0, 0, 0, 0);
                if (!pick) d1[n16] = make_uint4(0, 0, 0, 0);
            }
        } else {
            // Unaligned rows: whole 16-byte chunks from the aligned address below the row to the one above its end, i.e. up to 15
            // bytes of the NEIGHBOURING rows on either side -- allowed wherever that stays inside [safe_lo, safe_hi), the bytes of
            // this launch's frames (bytes past a row only ever meet zero weights); the first and last rows of the batch, where it
            // would not, are staged exactly.
            auto stage = [&](uint8_t *slot, const uint8_t *g) {
                const int lead = (int)(reinterpret_cast<uintptr_t>(g) & 15), nch = (lead + row_bytes + 15) >> 4;
                const uint8_t *a = g - lead;
                if (a >= safe_lo && a + 16 * nch <= safe_hi) {
                    for (int i = tid; i <= nch; i += QUAD_THREADS) {
                        if (i < nch) cp_async_16(slot + 16 * i, a + 16 * i);
                        else *reinterpret_cast<uint4 *>(slot + 16 * i) = make_uint4(0, 0, 0, 0);
                    }
                } else {
                    quads_stage_row(slot, g, row_bytes, tid);
                }
            };
            stage(smem + (SPR * r) * row_pad, g0);
            if (!pick && two) stage(smem + (SPR * r + 1) * row_pad, g1);
            if (tid == 0) {
                s_lead[2 * r] = (int)(reinterpret_cast<uintptr_t>(g0) & 15);
                s_lead[2 * r + 1] = two ? row_pad + (int)(reinterpret_cast<uintptr_t>(g1) & 15) : s_lead[2 * r];   // weight 0: any staged row will do
            }
        }
    }
    if (OUT == 0)
        for (int i = tid; i < 256; i += QUAD_THREADS) s_lut[i] = __fdiv_rn((float)i, 255.f);      // true float32 division, as the reference does
    cp_async_wait_all();
    __syncthreads();
    const int quads = plan.dst_w >> 2;
    for (int item = tid; item < 2 * quads; item += QUAD_THREADS) {
        const int half = item >= quads ? 1 : 0, j = item - half * quads;
        const int4 ta = __ldg(reinterpret_cast<const int4 *>(plan.xpack) + 2 * j), tb = __ldg(reinterpret_cast<const int4 *>(plan.xpack) + 2 * j + 1);
        const int off[4] = {ta.x, ta.z, tb.x, tb.z};
        const uint32_t aw[4] = {(uint32_t)ta.y, (uint32_t)ta.w, (uint32_t)tb.y, (uint32_t)tb.w};
#pragma unroll
        for (int rr = 0; rr < QUAD_ROWS; rr += 2) {
            const int r = rr + half, y = y_first + r;
            if (y >= plan.dst_h) break;
            int yb0 = 0, yb1 = 0;
            if (MODE == 0) { yb0 = __ldg(plan.b0 + y); yb1 = __ldg(plan.b1 + y); }
            const uint8_t *s_slot = smem + (SPR * r) * row_pad;
            // ALIGNED: the rows sit at the start of their slots (row 1 in the next slot; with weight 0 any staged row will do)
            const int lead0 = ALIGNED ? 0 : s_lead[2 * r], lead1 = ALIGNED ? ((area || yb1 != 0) ? row_pad : 0) : s_lead[2 * r + 1];
            uint32_t px[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int o0 = lead0 + off[k], o1 = lead1 + off[k];
                const uint32_t sh0 = (uint32_t)(o0 & 3) * 8, sh1 = ALIGNED ? sh0 : (uint32_t)(o1 & 3) * 8;     // (slots are 16 bytes apart)
                const uint32_t *w0 = reinterpret_cast<const uint32_t *>(s_slot + (o0 & ~3));
                const uint32_t *w1 = ALIGNED ? reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(w0) + lead1)
                                             : reinterpret_cast<const uint32_t *>(s_slot + (o1 & ~3));
                if (pick) {                // the pixel itself: three bytes out of two words
                    px[k] = __funnelshift_r(w0[0], w0[1], sh0) & 0x00FFFFFFu;
                    continue;
                }
                const uint32_t lo0 = __funnelshift_r(w0[0], w0[1], sh0), hi0 = __funnelshift_r(w0[1], w0[2], sh0);   // bytes 0..3, 4..7 of the pair
                const uint32_t lo1 = __funnelshift_r(w1[0], w1[1], sh1), hi1 = __funnelshift_r(w1[1], w1[2], sh1);
                uint32_t p = 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const uint32_t selc = c == 0 ? 0x7730u : (c == 1 ? 0x7741u : 0x7752u);          // (p[x0][c], p[x0+1][c]) in the low half
                    const int s0 = (int)__dp2a_lo(aw[k], __byte_perm(lo0, hi0, selc), 0u);
                    const int s1 = (int)__dp2a_lo(aw[k], __byte_perm(lo1, hi1, selc), 0u);
                    // bilinear: ((b0 (S0 >> 4)) >> 16) + ((b1 (S1 >> 4)) >> 16) + 2: the rounding 2 rides in the first product as 2 << 16
                    // (it cannot disturb the bits below); <= 255 by construction.  2x2 mean: the taps are (1, 1), S0 + S1 is the sum of four
                    const int v = area ? (s0 + s1 + 2) >> 2 : (((yb0 * (s0 >> 4) + 0x20000) >> 16) + ((yb1 * (s1 >> 4)) >> 16)) >> 2;
                    p |= (uint32_t)v << (8 * c);
                }
                px[k] = p;
            }
            if (OUT == 0) {
                const int64_t plane = (int64_t)plan.dst_h * plan.dst_w;
                float *o = reinterpret_cast<float *>(out) + (int64_t)b * 3 * plane + (int64_t)y * plan.dst_w + 4 * j;
#pragma unroll
                for (int c = 0; c < 3; ++c) {        // RGB planes: output channel 0 is source byte 2
                    const int shift = 8 * (2 - c);
                    reinterpret_cast<float4 *>(o + c * plane)[0] = make_float4(s_lut[(px[0] >> shift) & 0xff], s_lut[(px[1] >> shift) & 0xff],
                                                                              s_lut[(px[2] >> shift) & 0xff], s_lut[(px[3] >> shift) & 0xff]);
                }
            } else {
                uint32_t *o = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(out) + (((int64_t)b * plan.dst_h + y) * plan.dst_w + 4 * j) * 3);
                o[0] = px[0] | (px[1] << 24);
                o[1] = (px[1] >> 8) | (px[2] << 16);
                o[2] = (px[2] >> 16) | (px[3] << 8);
            }
        }
    }
}

int g_k1_kernel = 0;        // cutdet_debug_k1_kernel: 0 = choose, 1 = one thread per pixel, 2 = row kernel, 3 = quad kernel (measurement aid)

int check_frames(const cutdet_resize_plan *plan, const cutdet_frames *src) {
    CUTDET_REQUIRE(plan && src, "preprocess: null plan/frames");
    CUTDET_REQUIRE(src->batch >= 0, "preprocess: negative batch");
    if (src->batch == 0) return CUTDET_OK;
    CUTDET_REQUIRE(src->frames_dev, "preprocess: null frames pointer");
    CUTDET_REQUIRE(src->row_pitch >= 3 * (int64_t)plan->host.src_w, "preprocess: row_pitch %lld < 3*width", (long long)src->row_pitch);
    const int rows = src->row_map_compact ? plan->n_rows : plan->host.src_h;
    CUTDET_REQUIRE(src->batch <= 1 || src->frame_stride >= src->row_pitch * (int64_t)(rows - 1) + 3 * (int64_t)plan->host.src_w,
                   "preprocess: frame_stride too small");
    return CUTDET_OK;
}

// Where the quad kernel is the default: every two-tap resize; gathers and copies per measurement (profiles/r02_k1_matrix_final.txt)
#define QUADS_BY_DEFAULT(one_pixel, out, row_bytes) (!(one_pixel) || (row_bytes) <= 6144)

template <int OUT>
static int launch_generic(const cutdet_resize_plan *plan, const cutdet_frames *src, void *out, cutdet_stream_t stream) {
    if (int rc = check_frames(plan, src)) return rc;
    if (src->batch == 0) return CUTDET_OK;
    CUTDET_REQUIRE(out, "preprocess: null output");
    // row kernel: 128-bit loads need 16-byte aligned rows
    const ResizePlanDev &h = plan->host;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src->frames_dev) | (uintptr_t)src->frame_stride | (uintptr_t)src->row_pitch) & 15) == 0;
    const int row_pad = ((3 * h.src_w + 15) / 16 + 1) * 16;
    const int out_bytes = OUT == 0 ? 3 * h.dst_w * 4 : (3 * h.dst_w + 15) / 16 * 16;
    const size_t smem = (size_t)ROWS_PER_BLOCK * (2 * (size_t)row_pad + out_bytes);
    // Which kernel (profiles/r02_k1_matrix_final.txt, B200, frames resident in HBM): the quad kernel wherever it applies, except for
    // gathers from rows longer than 6 KB (2160p: 88 % of the HBM peak against 95 % for one thread per pixel).  Where it does not
    // apply (an output width that is not a multiple of four, a misaligned output) the row kernel takes the uint8 output of aligned
    // frames with rows up to 6 KB and the one-thread-per-pixel kernel the rest, as before the quad kernel existed.
    const bool choose_rows = g_k1_kernel == 2 || (g_k1_kernel == 0 && OUT == 1 && 3 * h.src_w <= 6144);
    const int quad_pad = row_pad + 32;                   // a row starts up to 15 bytes into its slot; one chunk of zeros behind it
    const bool one_pixel = h.gather_step_x > 0 || h.mode == RESIZE_COPY;        // a gather or a copy: no arithmetic
    const int quad_mode = one_pixel ? 2 : (h.mode == RESIZE_AREA2 ? 1 : 0);
    const size_t smem_quads = (one_pixel ? 1 : 2) * (size_t)QUAD_ROWS * quad_pad + 1024 + 2 * QUAD_ROWS * sizeof(int);
    const bool quads_ok = h.dst_w % 4 == 0 && smem_quads <= 200 * 1024 && reinterpret_cast<uintptr_t>(out) % (OUT == 0 ? 16 : 4) == 0;
    const bool choose_quads = quads_ok && (g_k1_kernel == 3 || (g_k1_kernel == 0 && QUADS_BY_DEFAULT(one_pixel, OUT, 3 * h.src_w)));
    if (choose_quads) {
        using QuadsFn = void (*)(ResizePlanDev, const uint8_t *, int64_t, int64_t, int, void *, int, const uint8_t *, const uint8_t *);
        static const QuadsFn fns[3][2] = {{preprocess_quads_kernel<OUT, 0, false>, preprocess_quads_kernel<OUT, 0, true>},
                                          {preprocess_quads_kernel<OUT, 1, false>, preprocess_quads_kernel<OUT, 1, true>},
                                          {preprocess_quads_kernel<OUT, 2, false>, preprocess_quads_kernel<OUT, 2, true>}};
        const QuadsFn quads_fn = fns[quad_mode][aligned];
        static bool attr_set[2][3][2] = {};
        if (smem_quads > 48 * 1024 && !attr_set[OUT][quad_mode][aligned]) {
            CUTDET_CUDA(cudaFuncSetAttribute(quads_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[OUT][quad_mode][aligned] = true;
        }
        const int64_t out_frame = (int64_t)h.dst_h * h.dst_w * 3;
        for (int b0 = 0; b0 < src->batch; b0 += 65535) {
            const int nb = src->batch - b0 < 65535 ? src->batch - b0 : 65535;
            void *o = OUT == 0 ? (void *)((float *)out + b0 * out_frame) : (void *)((uint8_t *)out + b0 * out_frame);
            const uint8_t *first = src->frames_dev + (int64_t)b0 * src->frame_stride;
            const int buf_rows = src->row_map_compact ? plan->n_rows : h.src_h;
            const uint8_t *past = first + (int64_t)(nb - 1) * src->frame_stride + (int64_t)(buf_rows - 1) * src->row_pitch + 3 * (int64_t)h.src_w;
            {
                KernelScope scope("preprocess_quads_kernel", as_stream(stream));
                quads_fn<<<dim3((unsigned)ceil_div(h.dst_h, QUAD_ROWS), (unsigned)nb), QUAD_THREADS, smem_quads, as_stream(stream)>>>(
                    h, first, src->frame_stride, src->row_pitch, src->row_map_compact, o, quad_pad, first, past);
            }
            CUTDET_LAUNCH_CHECK("preprocess_quads_kernel");
        }
        return CUTDET_OK;
    }
    if (aligned && smem <= 200 * 1024 && h.dst_h <= 65535 && choose_rows) {
        static bool attr_set[2] = {false, false};
        if (smem > 48 * 1024 && !attr_set[OUT]) {
            CUTDET_CUDA(cudaFuncSetAttribute(preprocess_rows_kernel<OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[OUT] = true;
        }
        const int64_t out_frame = (int64_t)h.dst_h * h.dst_w * 3;
        const bool vec_out = OUT == 0 ? (h.dst_w % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0)
                                      : ((3 * h.dst_w) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0);
        for (int b0 = 0; b0 < src->batch; b0 += 65535) {
            const int nb = src->batch - b0 < 65535 ? src->batch - b0 : 65535;
            void *o = OUT == 0 ? (void *)((float *)out + b0 * out_frame) : (void *)((uint8_t *)out + b0 * out_frame);
            {
                KernelScope scope("preprocess_rows_kernel", as_stream(stream));
                preprocess_rows_kernel<OUT><<<dim3((unsigned)ceil_div(h.dst_h, ROWS_PER_BLOCK), (unsigned)nb), ROWS_THREADS, smem, as_stream(stream)>>>(
                    h, src->frames_dev + (int64_t)b0 * src->frame_stride, src->frame_stride, src->row_pitch, src->row_map_compact, o,
                    row_pad, out_bytes, vec_out ? 1 : 0);
            }
            CUTDET_LAUNCH_CHECK("preprocess_rows_kernel");
        }
        return CUTDET_OK;
    }
    dim3 block(64, 4, 1);
    for (int b0 = 0; b0 < src->batch; b0 += 65535) {
        const int nb = src->batch - b0 < 65535 ? src->batch - b0 : 65535;
        dim3 grid((unsigned)ceil_div(plan->host.dst_w, 64), (unsigned)ceil_div(plan->host.dst_h, 4), (unsigned)nb);
        const int64_t out_frame = (int64_t)plan->host.dst_h * plan->host.dst_w * 3;
        void *o = OUT == 0 ? (void *)((float *)out + b0 * out_frame) : (void *)((uint8_t *)out + b0 * out_frame);
        {
            KernelScope scope("preprocess_generic_kernel", as_stream(stream));
            preprocess_generic_kernel<OUT><<<grid, block, 0, as_stream(stream)>>>(
            plan->host, src->frames_dev + (int64_t)b0 * src->frame_stride, src->frame_stride, src->row_pitch,
            src->row_map_compact, o);
        }
        CUTDET_LAUNCH_CHECK("preprocess_generic_kernel");
    }
    return CUTDET_OK;
}

}  // namespace cutdet

using namespace cutdet;

extern "C" int cutdet_resize_plan_create(int src_h, int src_w, int dst_h, int dst_w, cutdet_resize_plan **out) {
    CUTDET_REQUIRE(out, "resize_plan_create: null output");
    CUTDET_REQUIRE(src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0, "resize_plan_create: bad geometry %dx%d -> %dx%d",
                   src_w, src_h, dst_w, dst_h);
    cutdet_resize_plan *plan = new cutdet_resize_plan();
    ResizePlanDev &h = plan->host;
    h.src_h = src_h; h.src_w = src_w; h.dst_h = dst_h; h.dst_w = dst_w;
    std::vector<int> x0, x1, a0, a1, y0, y1, b0, b1;
    std::vector<char> used(src_h, 0);
    if (src_h == dst_h && src_w == dst_w) {
        h.mode = RESIZE_COPY;
        for (int y = 0; y < src_h; ++y) used[y] = 1;
    } else if (src_w == 2 * dst_w && src_h == 2 * dst_h) {
        h.mode = RESIZE_AREA2;
        for (int y = 0; y < src_h; ++y) used[y] = 1;
    } else {
        h.mode = RESIZE_LINEAR;
        linear_taps(src_w, dst_w, true, x0, x1, a0, a1);
        linear_taps(src_h, dst_h, false, y0, y1, b0, b1);
        for (int y = 0; y < dst_h; ++y) {
            used[y0[y]] = 1;
            if (b1[y] != 0) used[y1[y]] = 1;
        }
    }
    h.all_a1_zero = 1; h.all_b1_zero = 1;
    for (size_t i = 0; i < a1.size(); ++i) if (a1[i] != 0) h.all_a1_zero = 0;
    for (size_t i = 0; i < b1.size(); ++i) if (b1[i] != 0) h.all_b1_zero = 0;
    std::vector<int> slot(src_h, -1);
    for (int y = 0; y < src_h; ++y)
        if (used[y]) { slot[y] = (int)plan->rows.size(); plan->rows.push_back(y); }
    if (h.mode == RESIZE_AREA2) plan->pair_rows = true;
    if (h.mode == RESIZE_LINEAR) {
        plan->pair_rows = true;
        for (int y = 0; y < dst_h && plan->pair_rows; ++y)
            if (b1[y] != 0 && (y1[y] != y0[y] + 1 || slot[y1[y]] != slot[y0[y]] + 1)) plan->pair_rows = false;
    }
    // rows never read still need a defined slot (never dereferenced)
    for (int y = 0; y < src_h; ++y) if (slot[y] < 0) slot[y] = 0;
    plan->n_rows = (int)plan->rows.size();

    // one device allocation for all tables
    const size_t n_xpack = ((2 * (size_t)dst_w + 3) / 4) * 4;       // first in the blob: read with 16-byte loads
    const size_t n_ints = n_xpack + 4 * (size_t)dst_w + 4 * (size_t)dst_h + (size_t)src_h;
    std::vector<int> blob(n_ints, 0);
    int *p = blob.data();
    auto put = [&](const std::vector<int> &v, size_t n) { if (!v.empty()) memcpy(p, v.data(), n * sizeof(int)); int *r = p; p += n; return r; };
    if (h.mode == RESIZE_LINEAR)
        for (int x = 0; x < dst_w; ++x) {
            int wa = a0[x], wb = a1[x];
            if (x1[x] != x0[x] + 1) { wa += wb; wb = 0; }          // clamped at the edge: both taps are the same pixel
            blob[2 * x] = 3 * x0[x];
            blob[2 * x + 1] = (int)((uint32_t)wa | ((uint32_t)wb << 16));
        }
    if (h.mode == RESIZE_AREA2)
        for (int x = 0; x < dst_w; ++x) { blob[2 * x] = 6 * x; blob[2 * x + 1] = 0x00010001; }      // pixels 2x and 2x + 1, weights (1, 1)
    if (h.mode == RESIZE_COPY)
        for (int x = 0; x < dst_w; ++x) { blob[2 * x] = 3 * x; blob[2 * x + 1] = 2048; }
    p += n_xpack;
    int *hx0 = put(x0, dst_w), *hx1 = put(x1, dst_w), *ha0 = put(a0, dst_w), *ha1 = put(a1, dst_w);
    int *hy0 = put(y0, dst_h), *hy1 = put(y1, dst_h), *hb0 = put(b0, dst_h), *hb1 = put(b1, dst_h);
    int *hslot = put(slot, src_h);
    cudaError_t e = cudaMalloc(&plan->dev_blob, n_ints * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(plan->dev_blob, blob.data(), n_ints * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();     // the tables may be used at once on a non-blocking stream (see upload_bytes, conv_tc.cu)
    if (e != cudaSuccess) {
        if (plan->dev_blob) cudaFree(plan->dev_blob);
        delete plan;
        return cuda_fail(e, "resize_plan_create upload");
    }
    int *d = reinterpret_cast<int *>(plan->dev_blob);
    h.x0 = d + (hx0 - blob.data()); h.x1 = d + (hx1 - blob.data());
    h.a0 = d + (ha0 - blob.data()); h.a1 = d + (ha1 - blob.data());
    h.y0 = d + (hy0 - blob.data()); h.y1 = d + (hy1 - blob.data());
    h.b0 = d + (hb0 - blob.data()); h.b1 = d + (hb1 - blob.data());
    h.row_slot = d + (hslot - blob.data());
    h.xpack = reinterpret_cast<const int2 *>(d);
    // integer-scale gather fast path: every second tap has zero weight and taps are evenly spaced
    h.gather_step_x = 0; h.gather_step_y = 0; h.gather_off_x = 0; h.gather_off_y = 0;
    if (h.mode == RESIZE_LINEAR && h.all_a1_zero && h.all_b1_zero && dst_w > 1 && dst_h > 1) {
        const int sx = x0[1] - x0[0], sy = y0[1] - y0[0];
        bool ok = sx > 0 && sy > 0;
        for (int x = 0; ok && x < dst_w; ++x) ok = (x0[x] == x0[0] + x * sx) && a0[x] == 2048;
        for (int y = 0; ok && y < dst_h; ++y) ok = (y0[y] == y0[0] + y * sy) && b0[y] == 2048;
        if (ok) { h.gather_step_x = sx; h.gather_step_y = sy; h.gather_off_x = x0[0]; h.gather_off_y = y0[0]; }
    }
    *out = plan;
    return CUTDET_OK;
}

extern "C" void cutdet_resize_plan_destroy(cutdet_resize_plan *plan) {
    if (!plan) return;
    if (plan->dev_blob) cudaFree(plan->dev_blob);
    delete plan;
}

extern "C" int cutdet_resize_rows(int src_h, int src_w, int dst_h, int dst_w, int *rows_host, int *n_rows_out) {
    CUTDET_REQUIRE(n_rows_out, "resize_rows: null output");
    CUTDET_REQUIRE(src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0, "resize_rows: bad geometry %dx%d -> %dx%d", src_w, src_h, dst_w, dst_h);
    std::vector<char> used(src_h, 0);
    if ((src_h == dst_h && src_w == dst_w) || (src_w == 2 * dst_w && src_h == 2 * dst_h)) {
        for (int y = 0; y < src_h; ++y) used[y] = 1;
    } else {
        std::vector<int> y0, y1, b0, b1;
        linear_taps(src_h, dst_h, false, y0, y1, b0, b1);
        for (int y = 0; y < dst_h; ++y) {
            used[y0[y]] = 1;
            if (b1[y] != 0) used[y1[y]] = 1;
        }
    }
    int n = 0;
    for (int y = 0; y < src_h; ++y)
        if (used[y]) { if (rows_host) rows_host[n] = y; ++n; }
    *n_rows_out = n;
    return CUTDET_OK;
}

extern "C" int cutdet_resize_plan_rows(const cutdet_resize_plan *plan, int *rows_host, int *n_rows_out) {
    CUTDET_REQUIRE(plan && n_rows_out, "resize_plan_rows: null argument");
    *n_rows_out = plan->n_rows;
    if (rows_host) memcpy(rows_host, plan->rows.data(), plan->rows.size() * sizeof(int));
    return CUTDET_OK;
}

extern "C" int cutdet_upload_frames(const cutdet_resize_plan *plan, const uint8_t *frames_host, int batch,
                                    int64_t frame_stride, int64_t row_pitch, uint8_t *dst_dev, cutdet_stream_t stream,
                                    int64_t *bytes_copied) {
    CUTDET_REQUIRE(plan && batch >= 0, "upload_frames: bad argument");
    if (bytes_copied) *bytes_copied = 0;
    if (batch == 0) return CUTDET_OK;
    CUTDET_REQUIRE(frames_host && dst_dev, "upload_frames: null pointer");
    const int src_h = plan->host.src_h;
    const int64_t width = 3 * (int64_t)plan->host.src_w;
    CUTDET_REQUIRE(row_pitch >= width, "upload_frames: row_pitch < 3*width");
    // Smallest period p | src_h over which the set of needed rows repeats (5 at 720p -> 256x144: rows 5y+2).
    const std::vector<int> &rows = plan->rows;
    int period = src_h;
    for (int p = 1; p < src_h; ++p) {
        if (src_h % p) continue;
        const int reps = src_h / p;
        if (plan->n_rows % reps) continue;
        const int g = plan->n_rows / reps;
        bool ok = true;
        for (int i = 0; ok && i < plan->n_rows; ++i) ok = rows[i] == rows[i % g] + (i / g) * p && rows[i % g] < p;
        if (ok) { period = p; break; }
    }
    const int reps = src_h / period, g = plan->n_rows / reps;
    const bool frames_contiguous = frame_stride == (int64_t)src_h * row_pitch;
    const int n_calls = frames_contiguous ? 1 : batch;
    const int64_t height = frames_contiguous ? (int64_t)batch * reps : reps;
    for (int c = 0; c < n_calls; ++c)
        for (int j = 0; j < g; ++j) {
            const uint8_t *s = frames_host + (int64_t)c * frame_stride + (int64_t)rows[j] * row_pitch;
            uint8_t *d = dst_dev + (int64_t)c * plan->n_rows * width + (int64_t)j * width;
            CUTDET_CUDA(cudaMemcpy2DAsync(d, (size_t)g * width, s, (size_t)period * row_pitch, (size_t)width, (size_t)height,
                                          cudaMemcpyHostToDevice, as_stream(stream)));
        }
    if (bytes_copied) *bytes_copied = (int64_t)batch * plan->n_rows * width;
    return CUTDET_OK;
}

extern "C" int cutdet_debug_k1_kernel(int mode) {
    CUTDET_REQUIRE(mode >= 0 && mode <= 3, "debug_k1_kernel: mode 0 (choose), 1 (one thread per pixel), 2 (row kernel) or 3 (quad kernel)");
    g_k1_kernel = mode;
    return CUTDET_OK;
}

extern "C" int cutdet_preprocess_f32(const cutdet_resize_plan *plan, const cutdet_frames *src, float *out_nchw_dev,
                                     cutdet_stream_t stream) {
    return launch_generic<0>(plan, src, out_nchw_dev, stream);
}

extern "C" int cutdet_preprocess_u8(const cutdet_resize_plan *plan, const cutdet_frames *src, uint8_t *out_hwc_dev,
                                    cutdet_stream_t stream) {
    return launch_generic<1>(plan, src, out_hwc_dev, stream);
}
