// Tensor-core path of the classifier for sm_100a: tcgen05.mma with TMEM accumulators, operands fed by TMA
// (cp.async.bulk.tensor) or cp.async gathers, bias + ReLU + 3x3 max-pool + BatchNorm fused into the epilogue.
//
// Stands in for the three CNNLayers of FrameConvNet plus the AdaptiveAvgPool/first FCLayer (reference
// frameID/net.py:33-40, 122-133, 62-68) when the architecture is input_channels = 3, three conv layers and 32 or 48
// hidden channels (prod_net: 48; the contrastive encoder of learn_contrasts.py: 32).
//
// The idea that shapes everything: make the 3x3 max-pool THREAD-LOCAL.  A GEMM row (= a TMEM lane = one epilogue
// thread) is a POOLED output pixel; the nine conv outputs under it go to nine different TMEM column blocks
// D_j, j = (dy, dx), C fp32 columns each (9 * 48 = 432 of the 512 columns).  The epilogue thread reads its lane,
// takes the 9-way max per channel (ReLU and +bias commute with max), applies the BatchNorm affine AFTER the pool as the
// reference does, and writes C channels.  No shuffles, no shared-memory round trip, and the full-resolution
// activation (7 MB/frame in fp32 for layer 1) never exists anywhere.
//
// The accumulator is drained ROW BY ROW of the pool window: the MMAs of block row dy = 0 are issued (and committed)
// first, then dy = 1, then dy = 2, each block row with its own full/empty mbarrier pair.  The epilogue keeps a running
// max in registers and hands a block row back the moment it has been read, so the MMAs of the next tile's row dy
// overlap the epilogue of this tile's rows dy+1, dy+2: 432 columns behave like a three-deep accumulator ring.
// Eight epilogue warps (two per TMEM lane quarter, each taking half of the channels) hide the TMEM load latency.
//
//   conv1 (Cin = 3):   K1 writes the input "x-unfolded": for every image row and pooled column px the 5 input pixels
//       3px-1 .. 3px+3 (x3 channels, +1 pad = 16 halves = one UMMA K-chunk).  A row of the A operand is then five such
//       chunks (input rows 3py-1 .. 3py+3), gathered by cp.async.  For conv-row dy the MMA takes K-chunks dy..dy+2
//       against ONE B matrix [48 x 3C] that holds the taps of the three dx positions (zero where a tap falls outside):
//       9 MMAs of N = 3C per 128 pooled pixels.  Pixels are stored as v/256 (exact in fp16); 256/255 is folded into
//       the weights.  One CTA walks one frame (32 tiles at 256x144), frames round-robin over the CTAs.
//   conv2/conv3 (Cin = C): activations live "phase-split and flattened":
//           act[(y%3)*3 + x%3][c/8][frame*FP + (y/3)*PW + x/3][8 ch]  fp16,  PW = Win/3 + 1, QH = Hin/3 + 1, FP = QH*PW
//       with every entry that is not a real pixel equal to zero.  A GEMM row is the flattened position g of a pooled pixel
//       (Y, X) -> g = frame*FP + Y*PW + X; the input pixel (3Y+oy, 3X+ox) it needs (oy, ox in -1..3) is entry
//       g + sy*PW + sx of phase plane ((oy+3)%3, (ox+3)%3), sy/sx in {-1, 0, +1}: a plain OFFSET.  The zero column PW-1
//       and the zero row QH-1 double as the conv's zero padding of the next row / next frame, so the 25 shifted views of
//       a tile are 25 start addresses into ONE copy of the tile's nine planes (+-32 positions of halo) in shared memory.
//       The planes arrive by TMA 16 channels at a time (3 pipeline stages of 54 KB); positions that are padding
//       (X >= out_w, Y >= out_h) are GEMM rows whose results are dropped (9 % at 16x28, 25 % at 5x9).
//       The dx positions that share a view are adjacent column blocks, so they are ONE MMA of N = C * n_dx against a B
//       matrix stacked [kx=2 | kx=1 | kx=0]: 135 MMAs per 128 GEMM rows.
//
// Numerics: 16-bit operands, fp32 accumulation, fp32 epilogue, 16-bit inter-layer activations.  The operand format is
// fp16, not bf16: same tensor-core rate, 8x finer rounding (2^-12), and every value on this path is far inside fp16's
// range (pixels in [0,1), BatchNorm'd activations of order 10; the epilogue clamps to +-65504 regardless).  Measured
// against the fp32 reference the logits move by <= 0.015 with fp16 where bf16 moved them by up to 0.25 (an all-stripes
// frame, where weight rounding errors add coherently); tolerance stated in tests/test_gpu_net.py.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>

#include <map>
#include <vector>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace cutdet {

using namespace tc;

namespace {

constexpr bool kBf16 = false;           // operand format of the MMAs and of the stored activations (false = fp16)
constexpr float kPixelScale = kBf16 ? 255.f : 255.f / 256.f;    // layer-1 input = x * kPixelScale (u8 pixels stay exact)
constexpr float kW1Scale = kBf16 ? 1.f / 255.f : 256.f / 255.f; // ... and its weights absorb the inverse

// two floats -> packed 16-bit operands; fp16 saturates to +-65504 in the conversion itself (F2FP.SATFINITE)
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { return kBf16 ? pack_bf16x2(lo, hi) : pack_f16x2_sat(lo, hi); }

// Frames per pass through conv1/conv2: one frame per conv1 CTA on 148 SMs, and the layer-1 activations of a pass (59 MB) stay
// L2-resident until conv2 reads them.  Two frames per CTA (296) pay a CTA's set-up, pipeline fill and drain once per two frames,
// but programmatic dependent launch already hides most of that and the 118 MB of activations go through HBM: measured on one
// box, 80-step runs 2.132 / 2.136 M frames/s (148 / 296), 30-step runs 2.21 / 2.25 M (profiles/README.md, v9).  CUTDET_SUB_BATCH
// overrides it for such experiments.
constexpr int SUB_BATCH = 148;
constexpr int BATCHSTATS_MAX = 148;     // training-mode BatchNorm on the tensor-core path: one frame per CTA, everything resident
// Frames whose layer-2 maps are gathered for ONE launch of conv12_frames and ONE of conv3: 28 frames per CTA.  A 4,050-frame chunk is
// then one launch of each instead of four (the same 28 rounds, but conv3's four ramps and drains become one: 0.116 -> 0.077 ms per
// chunk, 2.53-2.55 -> 2.62-2.67 M frames/s in same-box 40-step runs); the layer-2 maps (52 KB per frame) go through HBM instead of
// staying in the L2, which conv3 does not notice.
constexpr int GROUP_FRAMES = 4144;
constexpr int TMEM_COLS = 512;
constexpr int MID_STAGES = 3;
constexpr int MID_WIN = 192;            // positions per (plane, channel group) in a stage: 32 halo + 128 + 32 halo
constexpr int MID_HALO = 32;
constexpr int MID_STAGE_BYTES = 9 * 2 * MID_WIN * 16;            // nine planes x 16 channels
constexpr int C1_STAGES = 4;
constexpr int C1_A_STAGE_BYTES = 10 * 128 * 16;   // 5 input rows x 2 x (128 rows x 16 B)
constexpr int EPI_WARPS = 8;

// ------------------------------------------------------------------------------------------------ geometry
struct Geom {
    int H, W, C, CG;
    int P1h, P1w, P2h, P2w, P3h, P3w;   // pooled map sizes after layers 1, 2, 3
    int Q1h, PW1, FP1;                  // phase-split layout of layer 1's output (conv2's input)
    int Q2h, PW2, FP2;                  // ... of layer 2's output (conv3's input)
    size_t xin_frame, act3_frame;       // bytes per frame
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Geom make_geom(int H, int W, int C) {
    Geom g;
    g.H = H; g.W = W; g.C = C; g.CG = C / 8;
    g.P1h = H / 3; g.P1w = W / 3;
    g.P2h = g.P1h / 3; g.P2w = g.P1w / 3;
    g.P3h = g.P2h / 3; g.P3w = g.P2w / 3;
    // layer 1's frames start on a tile boundary of layer 2's GEMM rows (128 positions): a frame is then a whole number of tiles, which
    // is what lets one CTA run both layers of a frame back to back (conv12_frames_kernel); the rows past Q1h * PW1 are zero padding
    g.Q1h = g.P1h / 3 + 1; g.PW1 = g.P1w / 3 + 1; g.FP1 = (int)align_up((size_t)g.Q1h * g.PW1, 128);
    g.Q2h = g.P2h / 3 + 1; g.PW2 = g.P2w / 3 + 1; g.FP2 = g.Q2h * g.PW2;
    g.xin_frame = (size_t)H * g.P1w * 32;
    g.act3_frame = (size_t)g.P3h * g.P3w * C * sizeof(float);
    return g;
}

// positions per (plane, channel group) of a phase-split buffer holding `frames` frames (TMA rows are 32 positions)
int gtot_for(int frames, int FP) { return (int)align_up((size_t)frames * FP, 32); }
size_t act_bytes(int CG, int gtot) { return (size_t)9 * CG * gtot * 16; }

// ------------------------------------------------------------------------------------------------ epilogue
// Where a pooled pixel goes.  mode 0: phase-split 16-bit (input layout of the next conv); mode 1: [frame][pixel][C] fp32.
struct OutSpec {
    void *ptr;
    int mode;
    int gtot;          // mode 0: positions per (plane, channel group)
    int PW, FP, QH;    // mode 0: the next layer's row pitch, frame pitch, rows per frame
    int frame0;        // first frame of this launch inside the buffer
    int out_h, out_w;  // pooled map size
};

template <int CH>
__device__ __forceinline__ void reg_fence(float (&v)[CH]) {
    // ties the registers to the preceding tcgen05.wait::ld so the compiler cannot hoist their uses above it
#pragma unroll
    for (int i = 0; i < CH; ++i) asm volatile("" : "+f"(v[i]));
}

template <int CH>
__device__ __forceinline__ void tmem_ld_ch(uint32_t taddr, float (&v)[CH]) {
    static_assert(CH == 16 || CH == 24, "channels per epilogue thread");
    float a[16];
    tmem_ld16(taddr, a);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = a[i];
    if (CH == 24) {
        float b[8];
        tmem_ld8(taddr + 16, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[16 + i] = b[i];
    }
}

// One tile's epilogue for this thread: its TMEM lane (a pooled pixel), CH = C/2 of the channels.  Block row dy of the
// accumulator (columns [3C*dy, 3C*dy + 3C) = the three dx positions) is read as soon as its MMAs have committed and is
// handed back right after the read.  Returns max over the 9 positions, + bias, ReLU, BatchNorm affine (the 16-bit store saturates).
template <int C>
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_thread /* lane quarter + this thread's first column */,
                                              uint64_t *acc_full, uint64_t *acc_empty, uint32_t acc_phase, int lane,
                                              const float *s_par /* [3][C]: bias, scale, shift */, int ch0,
                                              float (&run)[C / 2]) {
    constexpr int CH = C / 2;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        mbar_wait(&acc_full[dy], acc_phase);
        tc_fence_after_sync();
        float a[CH], b[CH], c[CH];
        tmem_ld_ch<CH>(tmem_thread + (3 * dy + 0) * C, a);
        tmem_ld_ch<CH>(tmem_thread + (3 * dy + 1) * C, b);
        tmem_ld_ch<CH>(tmem_thread + (3 * dy + 2) * C, c);
        tmem_ld_wait();
        reg_fence<CH>(a); reg_fence<CH>(b); reg_fence<CH>(c);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[dy]);           // block row dy may be overwritten by the next tile
        if (dy == 0) {
#pragma unroll
            for (int i = 0; i < CH; ++i) run[i] = fmaxf(fmaxf(a[i], b[i]), c[i]);
        } else {
#pragma unroll
            for (int i = 0; i < CH; ++i) run[i] = fmaxf(fmaxf(fmaxf(run[i], a[i]), b[i]), c[i]);
        }
    }
    const float4 *bias4 = reinterpret_cast<const float4 *>(s_par + ch0);
    const float4 *scale4 = reinterpret_cast<const float4 *>(s_par + C + ch0);
    const float4 *shift4 = reinterpret_cast<const float4 *>(s_par + 2 * C + ch0);
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
        const float4 b = bias4[i], s = scale4[i], t = shift4[i];
        run[4 * i + 0] = fmaf(fmaxf(run[4 * i + 0] + b.x, 0.f), s.x, t.x);
        run[4 * i + 1] = fmaf(fmaxf(run[4 * i + 1] + b.y, 0.f), s.y, t.y);
        run[4 * i + 2] = fmaf(fmaxf(run[4 * i + 2] + b.z, 0.f), s.z, t.z);
        run[4 * i + 3] = fmaf(fmaxf(run[4 * i + 3] + b.w, 0.f), s.w, t.w);
    }
}

// ---- software-pipelined variant for kernels whose epilogue is the critical path (conv1: 9 short MMAs per tile) ----
// A batch of tcgen05.ld followed by tcgen05.wait::ld costs ~250 cycles whatever its size (tools/tmem_bw.cu: 148 cycles for one
// x16 load, 255 for six), so the tile is read in as few batches as possible -- its three block rows -- and the batch of the
// next block row is in flight while the running max takes in the current one: two buffers of 3 * CH registers plus the
// running max, 7 * CH = 168 at C = 48, which is why the epilogue warps raise their register budget with setmaxnreg.
// Block row 0 of the NEXT tile is requested before the BatchNorm/store of this one.  Per channel: max3, max3 + max, max3 + max3
// with zero (FMNMX3): five operations, ReLU included.
// The accumulator columns of a block row are ordered [channel half][dx][CH channels] (the B operand's rows are packed to match),
// so the 3 * CH values a thread needs of a block row are CONTIGUOUS: two tcgen05.ld per block row (x64 + x8 at C = 48, x32 +
// x16 at C = 32) instead of six, and a third of the address set-up.
template <int CH>
struct EpiRow { float v[3 * CH]; };

template <int C>
__device__ __forceinline__ void epi_issue_row(uint32_t tmem_thread, int dy, EpiRow<C / 2> &q) {
    static_assert(C == 48 || C == 32, "channels");
    const uint32_t t = tmem_thread + dy * 3 * C;
    if constexpr (C == 48) {
        float lo[64], hi[8];
        tmem_ld64(t, lo);
        tmem_ld8(t + 64, hi);
#pragma unroll
        for (int i = 0; i < 64; ++i) q.v[i] = lo[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) q.v[64 + i] = hi[i];
    } else {
        float lo[32], hi[16];
        tmem_ld32(t, lo);
        tmem_ld16(t + 32, hi);
#pragma unroll
        for (int i = 0; i < 32; ++i) q.v[i] = lo[i];
#pragma unroll
        for (int i = 0; i < 16; ++i) q.v[32 + i] = hi[i];
    }
}

// X: block row 0 of this tile, in flight on entry.  Y: free on entry; on exit it holds the next tile's block row 0 in flight.
// The MMAs run well ahead of this code, so the full-barriers of rows 1 and 2 (and of the next tile's row 0) have normally
// completed long before they are needed: they are polled EARLY, back to back, and the ~130-cycle round trip of a try_wait hides
// behind the TMEM loads in flight; only a failed poll turns into a blocking wait.
template <int C>
__device__ __forceinline__ void epilogue_tile_pipelined(uint32_t tmem_thread, uint64_t *acc_full, uint64_t *acc_empty, uint32_t acc_phase,
                                                        int lane, bool more_tiles, EpiRow<C / 2> &X, EpiRow<C / 2> &Y,
                                                        float (&run)[C / 2]) {
    constexpr int CH = C / 2;
    auto arrived = [&](EpiRow<CH> &q, int dy) {       // block row dy is in registers: hand it back to the MMA issuer
        tmem_ld_wait();
        reg_fence<3 * CH>(q.v);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[dy]);
    };
    const bool f1 = mbar_try_wait(&acc_full[1], acc_phase), f2 = mbar_try_wait(&acc_full[2], acc_phase);
    arrived(X, 0);
    if (!f1) mbar_wait(&acc_full[1], acc_phase);
    tc_fence_after_sync();
    epi_issue_row<C>(tmem_thread, 1, Y);
#pragma unroll
    for (int i = 0; i < CH; ++i) run[i] = fmaxf(fmaxf(X.v[i], X.v[CH + i]), X.v[2 * CH + i]);
    reg_fence<CH>(run);
    arrived(Y, 1);
    if (!f2) mbar_wait(&acc_full[2], acc_phase);
    tc_fence_after_sync();
    epi_issue_row<C>(tmem_thread, 2, X);
    const bool f0 = more_tiles && mbar_try_wait(&acc_full[0], acc_phase ^ 1);
#pragma unroll
    for (int i = 0; i < CH; ++i) run[i] = fmaxf(fmaxf(fmaxf(run[i], Y.v[i]), Y.v[CH + i]), Y.v[2 * CH + i]);
    reg_fence<CH>(run);
    arrived(X, 2);
    if (more_tiles) {
        if (!f0) mbar_wait(&acc_full[0], acc_phase ^ 1);
        tc_fence_after_sync();
        epi_issue_row<C>(tmem_thread, 0, Y);
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const float m = fmaxf(fmaxf(run[i], X.v[i]), X.v[CH + i]);
        run[i] = fmaxf(fmaxf(m, X.v[2 * CH + i]), 0.f);  // the ReLU rides in the last max3 (the bias is inside the accumulator)
    }
    reg_fence<CH>(run);
}

// run = relu(accumulator incl. bias) on entry: scale (+-1/256) and shift only
template <int C>
__device__ __forceinline__ void epilogue_scale_shift(const float *s_par, int ch0, float (&run)[C / 2]) {
    constexpr int CH = C / 2;
    const float4 *scale4 = reinterpret_cast<const float4 *>(s_par + C + ch0);
    const float4 *shift4 = reinterpret_cast<const float4 *>(s_par + 2 * C + ch0);
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
        const float4 s = scale4[i], t = shift4[i];
        run[4 * i + 0] = fmaf(run[4 * i + 0], s.x, t.x);
        run[4 * i + 1] = fmaf(run[4 * i + 1], s.y, t.y);
        run[4 * i + 2] = fmaf(run[4 * i + 2], s.z, t.z);
        run[4 * i + 3] = fmaf(run[4 * i + 3], s.w, t.w);
    }
}

// b = frame inside this launch, (Y, X) = pooled pixel, ch0 = this thread's first channel (a multiple of 8).
template <int C>
__device__ __forceinline__ void store_pixel(const OutSpec &o, int b, int Y, int X, int ch0, const float (&v)[C / 2]) {
    constexpr int CG = C / 8, CH = C / 2;
    if (o.mode == 0) {
        const int plane = (Y % 3) * 3 + (X % 3);
        uint4 *dst = reinterpret_cast<uint4 *>(o.ptr) +
                     ((size_t)(plane * CG + ch0 / 8) * o.gtot + (size_t)(o.frame0 + b) * o.FP + (Y / 3) * o.PW + X / 3);
#pragma unroll
        for (int j = 0; j < CH / 8; ++j) {
            uint4 q;
            q.x = pack2(v[8 * j + 0], v[8 * j + 1]); q.y = pack2(v[8 * j + 2], v[8 * j + 3]);
            q.z = pack2(v[8 * j + 4], v[8 * j + 5]); q.w = pack2(v[8 * j + 6], v[8 * j + 7]);
            dst[(size_t)j * o.gtot] = q;
        }
    } else {
        float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(o.ptr) +
                                                 ((size_t)(o.frame0 + b) * o.out_h * o.out_w + (size_t)Y * o.out_w + X) * C + ch0);
#pragma unroll
        for (int j = 0; j < CH / 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
}


// ---- the same pipeline on fp16 accumulators (conv1_fused_tc_kernel<.., ACC16 = true>) ----
// The MMA writes D as fp16, one value per 32-bit column; tcgen05.ld.pack::16b returns two adjacent columns per register, i.e. a
// half2 of two adjacent CHANNELS of one dx: half the TMEM read traffic, half the registers, and the 9-way max, the ReLU and the
// affine run on channel pairs (VHMNMX, HFMA2).  The result IS the packed activation that gets stored.
template <int CH>
struct EpiRow16 { uint32_t v[3 * CH / 2]; };

template <int N>
__device__ __forceinline__ void reg_fence_u(uint32_t (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(v[i]));
}

template <int C>
__device__ __forceinline__ void epi_issue_row16(uint32_t tmem_thread, int dy, EpiRow16<C / 2> &q) {
    static_assert(C == 48 || C == 32, "channels");
    const uint32_t t = tmem_thread + dy * 3 * C;
    if constexpr (C == 48) {            // 72 columns: 64 + 8
        tmem_ld_pack32(t, q.v);
        tmem_ld_pack4(t + 64, q.v + 32);
    } else {                            // 48 columns: 32 + 16
        tmem_ld_pack16(t, q.v);
        tmem_ld_pack8(t + 32, q.v + 16);
    }
}

// Returns whether the next tile's block row 0 has been requested (into Y).  Its MMAs can only start once this tile's row 0 has
// been handed back, so it is often NOT complete yet when this tile's reads are done: it is then polled without blocking
// (mbarrier.test_wait) and the caller tries again after the store -- blocking here would put the rest of this tile behind the
// next tile's first MMAs, and the MMA issuer behind that in turn.
template <int C>
__device__ __forceinline__ bool epilogue_tile_pipelined16(uint32_t tmem_thread, uint64_t *acc_full, uint64_t *acc_empty, uint32_t acc_phase,
                                                          int lane, bool more_tiles, EpiRow16<C / 2> &X, EpiRow16<C / 2> &Y,
                                                          uint32_t (&run)[C / 4], long long *stamps = nullptr) {
    constexpr int CP = C / 4;           // channel pairs per thread
    if (stamps) stamps[0] = clock64();
    auto arrived = [&](EpiRow16<C / 2> &q, int dy) {
        tmem_ld_wait();
        reg_fence_u<3 * CP>(q.v);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[dy]);
    };
    const bool f1 = mbar_test_wait(&acc_full[1], acc_phase), f2 = mbar_test_wait(&acc_full[2], acc_phase);
    arrived(X, 0);
    if (!f1) mbar_wait(&acc_full[1], acc_phase);
    tc_fence_after_sync();
    epi_issue_row16<C>(tmem_thread, 1, Y);
#pragma unroll
    for (int i = 0; i < CP; ++i) run[i] = hmax3(X.v[i], X.v[CP + i], X.v[2 * CP + i]);
    reg_fence_u<CP>(run);
    if (stamps) stamps[1] = clock64();
    arrived(Y, 1);
    if (stamps) stamps[2] = clock64();
    if (!f2) mbar_wait(&acc_full[2], acc_phase);
    tc_fence_after_sync();
    epi_issue_row16<C>(tmem_thread, 2, X);
#pragma unroll
    for (int i = 0; i < CP; ++i) run[i] = hmax2(hmax3(run[i], Y.v[i], Y.v[CP + i]), Y.v[2 * CP + i]);
    reg_fence_u<CP>(run);
    if (stamps) stamps[3] = clock64();
    const bool f0 = more_tiles && mbar_test_wait(&acc_full[0], acc_phase ^ 1);
    arrived(X, 2);
    if (stamps) stamps[4] = clock64();
    if (f0) {
        tc_fence_after_sync();
        epi_issue_row16<C>(tmem_thread, 0, Y);
    }
#pragma unroll
    for (int i = 0; i < CP; ++i) run[i] = hmax3(hmax3(run[i], X.v[i], X.v[CP + i]), X.v[2 * CP + i], 0u);   // ... and the ReLU
    reg_fence_u<CP>(run);
    if (stamps) stamps[5] = clock64();
    return f0 || !more_tiles;
}

// run = relu(accumulator incl. bias) as channel pairs: scale (+-2^k, exact) and BatchNorm shift, one HFMA2 per pair
template <int C>
__device__ __forceinline__ void epilogue_scale_shift16(const uint32_t *s_par16, int ch0, uint32_t (&run)[C / 4]) {
    constexpr int CP = C / 4;
    const uint4 *scale4 = reinterpret_cast<const uint4 *>(s_par16 + ch0 / 2);
    const uint4 *shift4 = reinterpret_cast<const uint4 *>(s_par16 + C / 2 + ch0 / 2);
#pragma unroll
    for (int i = 0; i < CP / 4; ++i) {
        const uint4 s = scale4[i], t = shift4[i];
        run[4 * i + 0] = hfma2(run[4 * i + 0], s.x, t.x);
        run[4 * i + 1] = hfma2(run[4 * i + 1], s.y, t.y);
        run[4 * i + 2] = hfma2(run[4 * i + 2], s.z, t.z);
        run[4 * i + 3] = hfma2(run[4 * i + 3], s.w, t.w);
    }
}

// phase-split store of channel pairs that are already packed (mode 0 only: conv1 always feeds conv2).  The address is computed at
// the START of the tile, where its chain of dependent integer operations hides behind the TMEM loads, not at the end.
template <int C>
__device__ __forceinline__ uint4 *store_addr16(const OutSpec &o, int b, int Y, int X, int ch0) {
    constexpr int CG = C / 8;
    const int plane = (Y % 3) * 3 + (X % 3);
    return reinterpret_cast<uint4 *>(o.ptr) + ((size_t)(plane * CG + ch0 / 8) * o.gtot + (size_t)(o.frame0 + b) * o.FP + (Y / 3) * o.PW + X / 3);
}
template <int C>
__device__ __forceinline__ void store_pixel16(uint4 *dst, int gtot, const uint32_t (&v)[C / 4]) {
#pragma unroll
    for (int j = 0; j < C / 16; ++j) dst[(size_t)j * gtot] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
template <int C>
__device__ __forceinline__ void store_pixel16_hint(uint4 *dst, int gtot, const uint32_t (&v)[C / 4], uint64_t policy) {
#pragma unroll
    for (int j = 0; j < C / 16; ++j) st_global_v4_hint(dst + (size_t)j * gtot, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], policy);
}

// Zero the entries of frames [f_lo, f_hi) of a phase-split buffer that are not real pixels (the last row and/or the
// last column of each plane, and the positions between the last row and the frame pitch): they are the zero padding the next
// conv's shifted views read.
__device__ __forceinline__ void zero_pads(const OutSpec &o, int CG, int f_lo, int f_hi, int tid, int nthreads) {
    if (o.mode != 0) return;
    const int gap = o.FP - o.QH * o.PW, span = o.PW + o.QH + gap, per_frame = 9 * CG * span;
    const long long total = (long long)(f_hi - f_lo) * per_frame;
    uint4 *dst = reinterpret_cast<uint4 *>(o.ptr);
    for (long long i = tid; i < total; i += nthreads) {
        const int f = f_lo + (int)(i / per_frame), r = (int)(i % per_frame);
        const int e = r % span, pc = r / span, plane = pc / CG;
        const int py = plane / 3, px = plane % 3;
        int Yq, Xq;
        bool pad;
        if (e < o.PW) { Yq = o.QH - 1; Xq = e; pad = 3 * Yq + py >= o.out_h; }
        else if (e < o.PW + o.QH) { Yq = e - o.PW; Xq = o.PW - 1; pad = 3 * Xq + px >= o.out_w; }
        else { Yq = o.QH; Xq = e - o.PW - o.QH; pad = true; }          // position QH * PW + Xq, up to the frame pitch
        if (pad) dst[(size_t)pc * o.gtot + (size_t)(o.frame0 + f) * o.FP + Yq * o.PW + Xq] = make_uint4(0, 0, 0, 0);
    }
}

// ------------------------------------------------------------------------------------------------ conv2 / conv3
struct MidParams {
    int n_frames;           // frames in this launch
    int FP, PW;             // layout of the INPUT (= GEMM row numbering): frame pitch, row pitch
    int out_h, out_w;       // pooled output size
    int n_tiles;            // ceil(n_frames * FP / 128)
    OutSpec out;
    const uint4 *w_packed;  // [ky][c/8][3C rows: kx=2 | kx=1 | kx=0][8] 16-bit
    const float *bias, *scale, *shift;
    long long *timeline;    // debug: clock64 stamps of CTA 0 (null = off)
};

template <int C>
struct MidSmem {
    static constexpr int CG = C / 8;
    static constexpr int W_BYTES = 3 * CG * 3 * C * 16;
    static constexpr int W_KY_BYTES = CG * 3 * C * 16;
    static constexpr int LBO_B = 3 * C * 16;
    static constexpr int total = W_BYTES + MID_STAGES * MID_STAGE_BYTES + 256 /* barriers */ + 3 * C * 4;
};

// 384 threads: warps 0..7 = epilogue (TMEM lane quarter = warp % 4, channel half = warp / 4), warp 8 = TMA producer,
// warps 9..11 = MMA issuers, one per block row dy (warp 9 also allocates TMEM).  Persistent: each CTA walks tiles blockIdx.x,
// +gridDim.x, ...  Three issuers because a block row's 15 MMAs per k-step sit behind ~300 instructions of descriptor set-up, which
// ONE warp issues about as fast as the tensor pipe retires the MMAs (3,100 cycles per k-step measured, 2,450 of MMA time): the
// rows write disjoint TMEM columns, each issuer commits its own row's barrier, and a stage is free once all three committed it.
constexpr int MID_THREADS = 384;
template <int C>
__global__ void __launch_bounds__(MID_THREADS, 1) conv_mid_tc_kernel(const __grid_constant__ CUtensorMap in_map, const MidParams p) {
    using S = MidSmem<C>;
    constexpr int CG = C / 8, KS = C / 16, CH = C / 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *s_w = smem;
    uint8_t *s_stage = smem + S::W_BYTES;
    uint8_t *s_tail = s_stage + MID_STAGES * MID_STAGE_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_tail);
    uint64_t *empty = full + MID_STAGES;
    uint64_t *acc_full = empty + MID_STAGES;
    uint64_t *acc_empty = acc_full + 3;
    uint64_t *w_full = acc_empty + 3;                      // the packed taps have landed (bulk copy issued by the producer warp)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_full + 1);
    float *s_par = reinterpret_cast<float *>(s_tail + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.timeline && blockIdx.x == 0 && threadIdx.x == 0) p.timeline[0] = clock64();
    if (p.timeline && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); p.timeline[128 + 4 * blockIdx.x] = g; }

    // one-time setup.  The 41 KB of taps come by bulk copy, issued by the producer warp together with its first stages, so the
    // pipeline fills while they are in flight (copying them with LDG/STS before the first TMA cost ~2 us per launch).
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_par[i] = p.bias[i]; s_par[C + i] = p.scale[i]; s_par[2 * C + i] = p.shift[i]; }
    if (threadIdx.x == 0) {
        for (int s = 0; s < MID_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 3); }
        for (int d = 0; d < 3; ++d) { mbar_init(&acc_full[d], 1); mbar_init(&acc_empty[d], EPI_WARPS); }
        mbar_init(w_full, 1);
        fence_barrier_init();
        tma_prefetch_desc(&in_map);
    }
    if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch (launch_mid): everything above ran while the kernel before this one was still draining; from
    // here on its output is read and buffers it may still be reading are written.
    grid_dep_launch();
    grid_dep_wait();

    if (warp == 8) {
        // ------------------------------------------------------------------ TMA producer (lane 0 issues)
        uint32_t stage = 0, phase = 0;
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, S::W_BYTES);
            for (int ky = 0; ky < 3; ++ky)
                bulk_load_1d(s_w + ky * S::W_KY_BYTES, reinterpret_cast<const uint8_t *>(p.w_packed) + ky * S::W_KY_BYTES, S::W_KY_BYTES, w_full);
        }
        __syncwarp();
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            for (int ks = 0; ks < KS; ++ks) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[stage], MID_STAGE_BYTES);
                    uint8_t *dst = s_stage + stage * MID_STAGE_BYTES;
                    for (int pl = 0; pl < 9; ++pl)
                        tma_load_4d(dst + pl * (2 * MID_WIN * 16), &in_map, &full[stage], 0, 4 * tile - 1, 2 * ks, pl);
                }
                __syncwarp();
                if (++stage == MID_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 9) {
        // ------------------------------------------------------------------ MMA issuers: warp 9 + dy takes block row dy
        const uint32_t w_addr = smem_u32(s_w), stage_addr = smem_u32(s_stage);
        long long *tl = (p.timeline && blockIdx.x == 0 && warp == 9) ? p.timeline : nullptr;
        auto issue = [&](auto dyc) {
            constexpr int dy = decltype(dyc)::value;
            uint32_t stage = 0, phase = 0, acc_phase = 0;
            int tli = 1;
            if (tl && lane == 0) tl[tli] = clock64();
            ++tli;
            bool first_stamp = true;
            mbar_wait(w_full, 0);
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int ks = 0; ks < KS; ++ks) {
                    mbar_wait(&full[stage], phase);
                    if (ks == 0) mbar_wait(&acc_empty[dy], acc_phase ^ 1);   // the epilogue must be done with this row of the previous tile
                    tc_fence_after_sync();
                    if (tl && lane == 0) tl[tli] = clock64();      // stage data present
                    ++tli;
                    if (dy == 0 && p.timeline && first_stamp && lane == 0) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); p.timeline[128 + 4 * blockIdx.x + 2] = g; }
                    first_stamp = false;
                    const uint32_t a_stage = stage_addr + stage * MID_STAGE_BYTES;
                    if (elect_one()) {
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const int oy = dy + ky - 1;
                            const int py = (oy + 3) % 3, sy = oy < 0 ? -1 : (oy > 2 ? 1 : 0);
#pragma unroll
                            for (int o = 0; o < 5; ++o) {
                                const int ox = (o == 0) ? 1 : (o == 1) ? 0 : (o == 2) ? 2 : (o == 3) ? -1 : 3;   // centre first
                                const int px = (ox + 3) % 3, sx = ox < 0 ? -1 : (ox > 2 ? 1 : 0);
                                const int dx_lo = ox - 1 < 0 ? 0 : ox - 1, dx_hi = ox + 1 > 2 ? 2 : ox + 1;
                                const int n_dx = dx_hi - dx_lo + 1, kx_start = ox + 1 - dx_lo;
                                const uint32_t idesc = instr_desc_16bit(128, C * n_dx, kBf16);
                                const uint32_t a_view = a_stage + (py * 3 + px) * (2 * MID_WIN * 16) +
                                                        (uint32_t)(MID_HALO + sy * p.PW + sx) * 16;
                                const uint32_t b_tap = w_addr + ky * S::W_KY_BYTES + 2 * ks * S::LBO_B + (2 - kx_start) * C * 16;
                                const uint64_t da = smem_desc(a_view, MID_WIN * 16, 128);
                                const uint64_t db = smem_desc(b_tap, S::LBO_B, 128);
                                umma_16bit(tmem_base + C * (dy * 3 + dx_lo), da, db, idesc, (ks == 0 && ky == 0 && ox == 1) ? 0u : 1u);
                            }
                        }
                        if (ks == KS - 1) umma_commit(&acc_full[dy]);
                        umma_commit(&empty[stage]);                // one of the three arrivals that free the stage
                    }
                    __syncwarp();
                    if (tl && lane == 0) tl[tli] = clock64();      // stage issued
                    ++tli;
                    if (++stage == MID_STAGES) { stage = 0; phase ^= 1; }
                }
                acc_phase ^= 1;
            }
        };
        if (warp == 9) issue(std::integral_constant<int, 0>{});
        else if (warp == 10) issue(std::integral_constant<int, 1>{});
        else issue(std::integral_constant<int, 2>{});
    } else {
        // ------------------------------------------------------------------ epilogue
        // pads of the output buffer first (they belong to no GEMM row)
        {
            const int per = (p.n_frames + gridDim.x - 1) / gridDim.x;
            const int f_lo = min(p.n_frames, (int)blockIdx.x * per), f_hi = min(p.n_frames, f_lo + per);
            zero_pads(p.out, CG, f_lo, f_hi, threadIdx.x, EPI_WARPS * 32);
        }
        const int q = warp & 3, half = warp >> 2, m = q * 32 + lane, ch0 = half * CH;
        const uint32_t tmem_thread = tmem_base + ((uint32_t)(q * 32) << 16) + ch0;
        uint32_t acc_phase = 0;
        const long long n_rows = (long long)p.n_frames * p.FP;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const long long g = (long long)tile * 128 + m;
            const int b = (int)(g / p.FP), pos = (int)(g % p.FP);
            const int Y = pos / p.PW, X = pos % p.PW;
            const bool valid = g < n_rows && Y < p.out_h && X < p.out_w;
            float v[CH];
            epilogue_tile<C>(tmem_thread, acc_full, acc_empty, acc_phase, lane, s_par, ch0, v);
            if (valid) store_pixel<C>(p.out, b, Y, X, ch0, v);
            acc_phase ^= 1;
            if (p.timeline && blockIdx.x == 0 && threadIdx.x == 0) p.timeline[64 + tile / gridDim.x] = clock64();   // tile stored
        }
    }
    if (p.timeline && blockIdx.x == 0 && threadIdx.x == 0) p.timeline[63] = clock64();
    if (p.timeline && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); p.timeline[128 + 4 * blockIdx.x + 1] = g; }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ conv1
struct Conv1Params {
    const uint4 *xin;       // [B][H][P1w][2 x 16 B]: x-unfolded input, v/256 scale
    int B, H, P1h, P1w;
    int tiles_per_frame;    // ceil(P1h * P1w / 128)
    OutSpec out;
    const uint4 *w_packed;  // [ky][half][3C rows: dx=0 | dx=1 | dx=2][8] 16-bit, taps scaled by kW1Scale
    const float *bias, *scale, *shift;
    long long *timeline;    // debug: clock64 stamps of CTA 0 (null = off)
    int folded;             // the taps carry |BN scale| and the bias row (the fast fused path needs it)
    const uint4 *w_perm;    // fused kernel: w_packed with the rows of each block ordered [channel half][dx][C/2]
    const float *scale_magic; // fused kernel, integer-scale gather: +-1/256 (bias and ReLU happen inside the MMA / the max, see tc_prepare)
    // fused kernel with fp16 accumulators (ACC16): taps scaled per channel by a power of two, pixels as fp16 subnormals, and the
    // epilogue's scale (+-2^k) and BatchNorm shift as half2 pairs [C/2 scales | C/2 shifts]
    const uint4 *w_perm16;
    const uint32_t *par16;
};

template <int C>
struct C1Smem {
    static constexpr int W_BYTES = 6 * 3 * C * 16;
    static constexpr int LBO_B = 3 * C * 16;
    static constexpr int total = W_BYTES + C1_STAGES * C1_A_STAGE_BYTES + 256 + 3 * C * 4;
};

// 416 threads: warps 0..7 = epilogue, warps 8..11 = cp.async gather producers (thread t builds GEMM row t), warp 12 =
// MMA issuer (+ TMEM alloc).  CTA b walks frames b, b + gridDim.x, ...; a tile is 128 consecutive pooled pixels of a frame.
template <int C>
__global__ void __launch_bounds__(416, 1) conv1_tc_kernel(const Conv1Params p) {
    using S = C1Smem<C>;
    constexpr int CG = C / 8, CH = C / 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *s_w = smem;
    uint8_t *s_stage = smem + S::W_BYTES;
    uint8_t *s_tail = s_stage + C1_STAGES * C1_A_STAGE_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_tail);
    uint64_t *empty = full + C1_STAGES;
    uint64_t *acc_full = empty + C1_STAGES;
    uint64_t *acc_empty = acc_full + 3;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 3);
    float *s_par = reinterpret_cast<float *>(s_tail + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < S::W_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(s_w)[i] = p.w_packed[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_par[i] = p.bias[i]; s_par[C + i] = p.scale[i]; s_par[2 * C + i] = p.shift[i]; }
    fence_proxy_async();
    if (threadIdx.x == 0) {
        for (int s = 0; s < C1_STAGES; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
        for (int d = 0; d < 3; ++d) { mbar_init(&acc_full[d], 1); mbar_init(&acc_empty[d], EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 12) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int per_frame = p.P1h * p.P1w;

    if (warp >= 8 && warp < 12) {
        // ------------------------------------------------------------------ gather producers
        const int m = threadIdx.x - 256;
        uint32_t stage = 0, phase = 0;
        for (int f = blockIdx.x; f < p.B; f += gridDim.x) {
            for (int t = 0; t < p.tiles_per_frame; ++t) {
                const int pix = t * 128 + m;
                const bool in_range = pix < per_frame;
                const int py = in_range ? pix / p.P1w : 0, px = in_range ? pix % p.P1w : 0;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t *dst = s_stage + stage * C1_A_STAGE_BYTES + m * 16;
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    const int row = 3 * py - 1 + r;
                    const bool ok = in_range && row >= 0 && row < p.H;
                    const uint4 *src = p.xin + (ok ? (((size_t)f * p.H + row) * p.P1w + px) * 2 : 0);
                    cp_async_16(dst + (2 * r) * 2048, src, ok ? 16u : 0u);
                    cp_async_16(dst + (2 * r + 1) * 2048, src + 1, ok ? 16u : 0u);
                }
                cp_async_arrive_noinc(&full[stage]);
                if (++stage == C1_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 12) {
        // ------------------------------------------------------------------ MMA issuer (lane 0 issues)
        uint32_t stage = 0, phase = 0, acc_phase = 0;
        const uint32_t w_addr = smem_u32(s_w), stage_addr = smem_u32(s_stage);
        const uint32_t idesc = instr_desc_16bit(128, 3 * C, kBf16);
        for (int f = blockIdx.x; f < p.B; f += gridDim.x) {
            for (int t = 0; t < p.tiles_per_frame; ++t) {
                mbar_wait(&full[stage], phase);
                fence_proxy_async();
                tc_fence_after_sync();
                const uint32_t a_stage = stage_addr + stage * C1_A_STAGE_BYTES;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    mbar_wait(&acc_empty[dy], acc_phase ^ 1);
                    tc_fence_after_sync();
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 3; ++ks) {
                            const uint64_t da = smem_desc(a_stage + 2 * (dy + ks) * 2048, 2048, 128);
                            const uint64_t db = smem_desc(w_addr + 2 * ks * S::LBO_B, S::LBO_B, 128);
                            umma_16bit(tmem_base + 3 * C * dy, da, db, idesc, ks > 0 ? 1u : 0u);
                        }
                        umma_commit(&acc_full[dy]);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(&empty[stage]);
                __syncwarp();
                if (++stage == C1_STAGES) { stage = 0; phase ^= 1; }
                acc_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3, half = warp >> 2, m = q * 32 + lane, ch0 = half * CH;
        const uint32_t tmem_thread = tmem_base + ((uint32_t)(q * 32) << 16) + ch0;
        uint32_t acc_phase = 0;
        for (int f = blockIdx.x; f < p.B; f += gridDim.x) {
            zero_pads(p.out, CG, f, f + 1, threadIdx.x, EPI_WARPS * 32);
            for (int t = 0; t < p.tiles_per_frame; ++t) {
                const int pix = t * 128 + m;
                const bool valid = pix < per_frame;
                const int Y = valid ? pix / p.P1w : 0, X = valid ? pix % p.P1w : 0;
                float v[CH];
                epilogue_tile<C>(tmem_thread, acc_full, acc_empty, acc_phase, lane, s_par, ch0, v);
                if (valid) store_pixel<C>(p.out, f, Y, X, ch0, v);
                acc_phase ^= 1;
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ K1 fused into conv1
// conv1 straight from the DECODED FRAMES: neither the resized image nor the x-unfolded input ever goes to memory, and
// the MMA reads its A operand where the resize wrote it.  One CTA walks one frame at a time.
//
//   loader (warp 13)     the source rows each resized row needs (1 for an integer-scale gather such as 720p -> 256x144,
//                        else 2) arrive by cp.async.bulk into a ring of slots with full/empty mbarriers -- 64 KB in flight
//                        per SM, enough to cover HBM latency x bandwidth, so the frame read streams at HBM speed;
//   unfold (warps 8..11) a lane takes one pooled column px of one resized row y: the five pixels 3px-1 .. 3px+3, bit-exact
//                        with cv2.resize(INTER_LINEAR), u8 -> fp16 v/256 through the 0x6400|v trick ((1024 + v) / 256 - 4:
//                        one PRMT + one HFMA2 per pair) = one 16-half K-chunk, stored as two 16-byte halves;
//   the ring             three sub-rings, one per y % 3, each [k-half][position][16 B] over the FLATTENED position
//                        g = R * P1w + px, R = frame_in_cta * (P1h + 1) + y / 3 (one all-zero row per frame doubles as the
//                        conv's zero padding above the next frame and below this one).  GEMM row m of tile t is position
//                        128 t + m, and its K-chunk for input row 3py - 1 + c is position 128 t + m + {-P1w, 0, 0, 0, +P1w}[c]
//                        of sub-ring {2, 0, 1, 2, 0}[c]: for a whole tile that is 128 CONSECUTIVE 16-byte entries, i.e. a
//                        K-major UMMA operand at a plain offset.  Positions wrap at FR_CAP; the first 128 are mirrored past
//                        the end so a window never straddles the wrap.
constexpr int FR_CAP = 1024;                 // positions per sub-ring (a power of two)
constexpr int FR_PLANE = (FR_CAP + 128) * 16;            // bytes of one k-half plane incl. the mirror
constexpr int FR_SUB = 2 * FR_PLANE;
// conv12_frames can run with a SMALLER operand ring (CAP positions, not necessarily a power of two) and hand the difference to the
// raw-row ring: where a resized row needs two long source rows (1080p: 11.5 KB per slot) the raw ring, not the operand ring, is
// what the unfold warps wait for.  A tile's views span 128 + 2 P1w <= 300 positions and the unfold warps work on eight rows at a
// time, so CAP >= 3 P1w + 126 + 8/3 P1w (deadlock-free, all eight warps busy): 768 is the smallest multiple of 256 that fits.
constexpr int FR_CAP_TWO_ROWS = 768;
template <int CAP>
struct FrRing {
    static constexpr int PLANE = (CAP + 128) * 16, SUB = 2 * PLANE;
    static constexpr int EXTRA_RAW = 3 * (FR_SUB - SUB);                  // bytes handed to the raw ring
    static constexpr uint32_t MAGIC = (uint32_t)(0x100000000ull / CAP) + 1u;
    static_assert(CAP % 128 == 0 && CAP <= FR_CAP && CAP >= 3 * ((256 + 2) / 3) + 126, "operand ring capacity");
    // x mod CAP for 0 <= x < 2^22 (positions of one launch stay far below)
    static __device__ __forceinline__ int wrap(int x) {
        if ((CAP & (CAP - 1)) == 0) return x & (CAP - 1);
        return x - CAP * (int)__umulhi((uint32_t)x, MAGIC);
    }
};
// raw-row ring: what is left of the 227 KB (24 rows of 720p, 8 row pairs of 1080p with four unfold warps; 22 / 7 with eight)
constexpr int raw_bytes(int unfold_warps) { return 92160 - (unfold_warps - 4) * 1056; }
constexpr int RAW_SLOTS_MAX = 24;
constexpr int F_MAX_DST = 256;
constexpr int F1_CMP_STRIDE = F_MAX_DST + 8;   // [0] = the pixel left of the image, [1 + x] = pixel x, zeros after

struct FusedSrc {
    ResizePlanDev plan;
    const uint8_t *frames;
    long long frame_stride, row_pitch;
    int compact;
    int n_src;        // raw rows per resized row
    int row_bytes;    // 3 * src_w, a multiple of 16
    int n_slots;      // raw ring slots of n_src * row_bytes each (2 .. RAW_SLOTS_MAX)
    int tma_shift;    // conv12_frames with a tensor map of the source rows: a slot holds 2^tma_shift rows, fetched by ONE TMA box; -1 = bulk copies
    int pair_rows;    // a resized row's two source rows are adjacent in the buffer (one two-row box serves n_src = 2)
    int buf_rows;     // rows per frame in the buffer (all source rows, or the compact ones)
    int l2_prefetch;  // conv12_frames' TMA loaders bring a slot's box into the L2 BEFORE they wait for the slot to come free
};

template <int C, int UW>
struct F1Smem {
    static constexpr int RAW_BYTES = raw_bytes(UW);
    static constexpr int W_BYTES = 6 * 3 * C * 16;
    static constexpr int LBO_B = 3 * C * 16;
    static constexpr int OFF_RING = W_BYTES;
    static constexpr int OFF_RAW = OFF_RING + 3 * FR_SUB;
    static constexpr int OFF_TAB = OFF_RAW + RAW_BYTES + 64;           // 64 bytes of slack: the gather reads one word past a row
    static constexpr int OFF_CMP = OFF_TAB + F_MAX_DST * (8 + 8 + 16);  // rowoff[256][2], yb[256][2], xtab[256] (int4)
    static constexpr int OFF_BAR = OFF_CMP + UW * F1_CMP_STRIDE * 4;   // one resized row per unfold warp (general resize)
    static constexpr int OFF_PAR = OFF_BAR + 1024;
    static constexpr int total = OFF_PAR + 3 * C * 4;
};

// One resized pixel x (inside the image) as a (B, G, R, 0) word, from the source rows staged in shared memory: bit-exact
// with cv2.resize(INTER_LINEAR) (reference frameID/data.py:197-222).
__device__ __forceinline__ uint32_t resized_word(const ResizePlanDev &plan, const uint8_t *q0, const uint8_t *q1, const int4 *s_xtab,
                                                 int b0, int b1, int x) {
    if (plan.gather_step_x > 0) {
        const int b = 3 * (plan.gather_off_x + x * plan.gather_step_x);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(q0 + (b & ~3));
        return __funnelshift_r(wp[0], wp[1], (b & 3) * 8) & 0x00FFFFFFu;
    }
    int v[3];
    if (plan.mode == RESIZE_COPY) {
        v[0] = q0[3 * x]; v[1] = q0[3 * x + 1]; v[2] = q0[3 * x + 2];
    } else if (plan.mode == RESIZE_AREA2) {
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (q0[6 * x + c] + q0[6 * x + 3 + c] + q1[6 * x + c] + q1[6 * x + 3 + c] + 2) >> 2;
    } else {
        // OpenCV's fixed-point bilinear: horizontal S = a0 p[x0] + a1 p[x0+1] (11-bit coefficients, one
        // DP2A per channel and row on the byte pair), vertical (((b0 (S0 >> 4)) >> 16) + ((b1 (S1 >> 4)) >> 16) + 2) >> 2
        const int4 xt = s_xtab[x];
        const uint32_t sh = (uint32_t)(xt.x & 3) * 8, aw = (uint32_t)xt.y;
        const uint32_t *r0 = reinterpret_cast<const uint32_t *>(q0 + (xt.x & ~3));
        const uint32_t *r1 = reinterpret_cast<const uint32_t *>(q1 + (xt.x & ~3));
        const uint32_t lo0 = __funnelshift_r(r0[0], r0[1], sh), hi0 = __funnelshift_r(r0[1], r0[2], sh);   // bytes 0..3, 4..7 of the pair
        const uint32_t lo1 = __funnelshift_r(r1[0], r1[1], sh), hi1 = __funnelshift_r(r1[1], r1[2], sh);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t selc = c == 0 ? 0x7730u : (c == 1 ? 0x7741u : 0x7752u);          // (p[x0][c], p[x0+1][c]) in the low half
            const int s0 = (int)__dp2a_lo(aw, __byte_perm(lo0, hi0, selc), 0u);
            const int s1 = (int)__dp2a_lo(aw, __byte_perm(lo1, hi1, selc), 0u);
            v[c] = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;           // <= 255 by construction
        }
    }
    return (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16);
}

// The bilinear case alone (no dispatch on the plan inside): the unfold warps keep F1_RESIZE_ILP of these in flight per lane.
#ifndef F1_RESIZE_ILP
#define F1_RESIZE_ILP 4
#endif
__device__ __forceinline__ uint32_t bilinear_word(const uint8_t *q0, const uint8_t *q1, const int4 *s_xtab, int b0, int b1, int x) {
    const int4 xt = s_xtab[x];
    const uint32_t sh = (uint32_t)(xt.x & 3) * 8, aw = (uint32_t)xt.y;
    const uint32_t *r0 = reinterpret_cast<const uint32_t *>(q0 + (xt.x & ~3));
    const uint32_t *r1 = reinterpret_cast<const uint32_t *>(q1 + (xt.x & ~3));
    const uint32_t lo0 = __funnelshift_r(r0[0], r0[1], sh), hi0 = __funnelshift_r(r0[1], r0[2], sh);
    const uint32_t lo1 = __funnelshift_r(r1[0], r1[1], sh), hi1 = __funnelshift_r(r1[1], r1[2], sh);
    uint32_t word = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t selc = c == 0 ? 0x7730u : (c == 1 ? 0x7741u : 0x7752u);
        const int s0 = (int)__dp2a_lo(aw, __byte_perm(lo0, hi0, selc), 0u);
        const int s1 = (int)__dp2a_lo(aw, __byte_perm(lo1, hi1, selc), 0u);
        // ((b0 (S0 >> 4)) >> 16) + ((b1 (S1 >> 4)) >> 16) + 2, the 2 riding in the first product as 2 << 16; <= 255 for weight pairs
        // that sum to 2049 at most, which OpenCV's always do (tests/test_kernel_invariants.py)
        word |= (uint32_t)((((b0 * (s0 >> 4) + 0x20000) >> 16) + ((b1 * (s1 >> 4)) >> 16)) >> 2) << (8 * c);
    }
    return word;
}

// Warps 0..7 = epilogue, then the unfold warps, the MMA issuer (+ TMEM alloc) and three loaders (F1Roles below: 640 threads with
// fp16 accumulators, 512 with fp32 ones).
//
// The four stages run DECOUPLED, each at its own pace, joined by two rings:
//   loaders   rows of the source frame -> raw ring by cp.async.bulk, row n issued by loader n % n_loaders (the slot count is a
//             multiple of the loader count, so a slot is always refilled by the same loader): ONE issuing thread sustains a
//             3,840-byte row per ~370 cycles (10 bytes/clock), two reach 5 TB/s (tools/hbm_rows.cu).  raw_full[slot]
//             counts the bytes, raw_empty[slot] the reader's release, s_rows_issued[loader] says how far each loader is:
//             a reader first checks that ITS row has been issued, because an mbarrier wait sees one parity bit and a reader
//             that is a whole ring ahead would otherwise take the previous row's phase for its own.
//   unfold    resized row u = 3R + sub goes to warp u % 4, which publishes its progress in s_rows_done[warp] (a monotone
//             counter, st.release after fence.proxy.async); it may run ahead of the MMAs by the capacity of the operand ring
//             and waits for tile_done[t % 8] (committed by the MMA issuer after tile t's last MMA) before overwriting positions
//             tile t still reads; the row it then writes is needed by tile t + 7 at the latest, so that barrier is never more
//             than one phase ahead of a waiter.
//             Pixels enter the MMA as the fp16 numbers 1024 + v (bit pattern 0x6400 | v, exact), so a chunk is a byte permutation
//             of five (B, G, R) words; its 16th element is the constant 1024, which multiplies a weight row holding the bias
//             (tc_prepare): acc = 256 * (|s| z + |s| bias) with |BatchNorm scale| folded into the taps, and the epilogue is
//             relu inside the last max3, then one FFMA by +-1/256 and the BatchNorm shift.
//             Integer-scale gathers (720p, 1440p, 2160p -> 256 wide; template GATHER) funnel-shift those words straight out of
//             the source row: per-lane word offsets and selectors are computed once (the byte phase of tap j is the same in all
//             three 32-column parts).  Every other resize first writes the resized row -- each pixel computed once, OpenCV's
//             fixed-point bilinear -- into a 1 KB per-warp buffer and unfolds from there (five consecutive words per column).
// Counters instead of per-tile mbarriers: a row is produced by ONE warp, tiles need rows from all of them, and a warp must never
// have to wait for a tile it contributes nothing to (the lock-step version spent 2/3 of its time in such waits).
// GATHER: the resize is out[y][x] = src[off_y + y*step_y][off_x + x*step_x] (every second tap has zero weight).
// Registers: setmaxnreg works on four consecutive warps, and a group can only grow by what the others of the SAME CTA have given
// up (a request beyond the pool would block for ever: the launcher checks the compiled register count).
// With fp16 accumulators (ACC16) the epilogue needs half the registers, which pays for a second warpgroup of unfold warps: the
// unfold's per-row chain of barrier polls, shared-memory loads and stores (~1,600 cycles per row and warp) set the tile period
// with four.  MMA_WARPS = 3 (one issuer per block row; a block row's issue -- barrier wait, three descriptors moved to uniform
// registers, three UTCHMMA, commit -- costs a warp 300-450 cycles) with six unfold warps measured the same at 720p and lower
// where the unfold computes the resize (360p 1.50 -> 1.37 M frames/s), so one issuer it stays.
// Warps: 0-7 epilogue | 8.. unfold | MMA issuer(s) | three loaders.  Registers are set per WARPGROUP (setmaxnreg.sync.aligned):
// the epilogue groups rise, the last group (issuer + loaders) drops to 48, the groups between take the unfold budget.
// ACC16: 640 threads x 96 = 256 x 128 + 256 x 88 + 128 x 48.   ACC32: 512 x 128 = 256 x 192 + 128 x 80 + 128 x 48.
constexpr int LOADER_WARPS = 3;
constexpr int TILE_RING = 8;                 // tile_done barriers; FR_CAP / 128 tiles of run-ahead at most
template <bool ACC16>
struct F1Roles {
    static constexpr int UNFOLD_WARPS = ACC16 ? 8 : 4, MMA_WARPS = 1;
    static constexpr int MMA_WARP0 = 8 + UNFOLD_WARPS, LOAD_WARP0 = MMA_WARP0 + MMA_WARPS;
    static constexpr int WARPS = LOAD_WARP0 + LOADER_WARPS, THREADS = 32 * WARPS;
    static constexpr int REGS_START = ACC16 ? 96 : 128, REGS_EPI = ACC16 ? 128 : 192, REGS_UNFOLD = ACC16 ? 88 : 80, REGS_LIGHT = 48;
    static_assert(WARPS % 4 == 0 && MMA_WARP0 + MMA_WARPS > WARPS - 4 && (MMA_WARPS == 1 || MMA_WARPS == 3), "whole warpgroups; an issuer in the last");
    static_assert(THREADS * REGS_START == 256 * REGS_EPI + (THREADS - 384) * REGS_UNFOLD + 128 * REGS_LIGHT, "register pool");
};
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ void st_release_shared(int *p, int v) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// after a fence that already ordered this thread's (and, through __syncwarp, its warp's) earlier writes: no second MEMBAR
__device__ __forceinline__ void st_relaxed_shared(int *p, int v) {
    asm volatile("st.relaxed.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_shared(const int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

// Sixteen halves of one K-chunk in the 1024 + v format from five source words (B, G, R, <any>): RGB of pixels 0..4, then the
// constant 1024.  Every half is the byte 0x64 over a pixel byte, so a word is one PRMT against the constant, or PRMT + LOP3
// where its halves come from two pixels.
// SUBN (the fp16-accumulator kernel): the halves are the fp16 SUBNORMALS v * 2^-24 (high byte zero: no offset to cancel,
// which an fp16 accumulator could not carry) and the constant is 1.0.
template <bool SUBN>
__device__ __forceinline__ void chunk_from_raw(const uint32_t (&w)[5], uint4 &lo, uint4 &hi) {
    constexpr uint32_t K = SUBN ? 0u : 0x64646464u, K0 = SUBN ? 0x3c000000u : 0x64006464u, M = 0x00ff00ffu, O = SUBN ? 0u : 0x64006400u;
    lo.x = __byte_perm(w[0], K, 0x4142);                        // R0 G0
    lo.y = (__byte_perm(w[0], w[1], 0x0600) & M) | O;           // B0 R1
    lo.z = __byte_perm(w[1], K, 0x4041);                        // G1 B1
    lo.w = __byte_perm(w[2], K, 0x4142);                        // R2 G2
    hi.x = (__byte_perm(w[2], w[3], 0x0600) & M) | O;           // B2 R3
    hi.y = __byte_perm(w[3], K, 0x4041);                        // G3 B3
    hi.z = __byte_perm(w[4], K, 0x4142);                        // R4 G4
    hi.w = __byte_perm(w[4], K0, 0x7640);                       // B4, then the constant 1024 that multiplies the bias row
}

// ---- the unfold role of the fused kernels (shared by conv1_fused_tc_kernel and conv1_fused_sets_kernel) ----
struct F1Ctx {
    const Conv1Params *p;
    const FusedSrc *src;
    uint8_t *s_ring, *s_raw;
    uint32_t *s_cmp;
    int *s_yb;
    int4 *s_xtab;
    uint64_t *raw_full, *raw_empty, *tile_done;
    int *s_rows_done, *s_rows_issued;
    int RPF, Hc, total_u, n_slots, slot_bytes, n_loaders;
    uint32_t inv_slots;
    long long *tl;
    int rps_shift = 0;   // a raw slot holds 2^rps_shift rows (the loaders fetch and the barriers count whole slots)
};

template <int C, bool GATHER, bool ACC16, int UNFOLD_WARPS, int CAP = FR_CAP>
__device__ __forceinline__ void f1_unfold_role(const F1Ctx &cx, const int pwarp, const int lane) {
    using RG = FrRing<CAP>;
    constexpr int NP = (F_MAX_DST / 3 + 31) / 32;
    const Conv1Params &p = *cx.p;
    const FusedSrc &src = *cx.src;
    const ResizePlanDev &plan = src.plan;
    uint8_t *const s_ring = cx.s_ring, *const s_raw = cx.s_raw;
    uint32_t *const s_cmp = cx.s_cmp;
    int *const s_yb = cx.s_yb;
    int4 *const s_xtab = cx.s_xtab;
    uint64_t *const raw_full = cx.raw_full, *const raw_empty = cx.raw_empty, *const tile_done = cx.tile_done;
    int *const s_rows_done = cx.s_rows_done, *const s_rows_issued = cx.s_rows_issued;
    const int H = p.H, P1w = p.P1w, RPF = cx.RPF, Hc = cx.Hc, total_u = cx.total_u, n_slots = cx.n_slots, slot_bytes = cx.slot_bytes;
    const uint32_t inv_slots = cx.inv_slots;
    long long *const tl = cx.tl;
    const int n_loaders = cx.n_loaders;
    // fast path constants: tap j of pooled column px starts at byte 3*off_x + BS*(3px - 1 + j), BS = 3*step_x; a part is
    // 32 columns = 96*BS bytes further on (a multiple of 4), so the word offset of part 0 and the byte phase serve all parts
    const int BS = 3 * plan.gather_step_x, PB = 96 * BS;
    int off0[5];
    uint32_t sel[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int b = 3 * plan.gather_off_x + BS * (3 * lane - 1 + j);
        off0[j] = b & ~3;
        sel[j] = 0x3210u + (uint32_t)(b & 3) * 0x1111u;
    }
    struct Row { int R, sub, slot; bool real; const uint8_t *q0; };
    int R = 0, sub = pwarp, fi = 0, py = 0;              // u = 3R + sub, R = fi * RPF + py
    while (sub >= 3) { sub -= 3; ++R; ++py; }
    int tiles_waited = 0, rows_done = 0;
    auto next_row = [&](Row &r, int &y_out) {            // describes the row (R, sub, fi, py) point at, waits until it may be
        while (py >= RPF) { py -= RPF; ++fi; }           // produced (ring space, source row landed), then advances
        const int y = 3 * py + sub;
        r.real = y < H && (py < p.P1h || sub == 0);      // index P1h: row 3*P1h if the image has it, else zeros
        const int n = fi * Hc + y, g = n >> cx.rps_shift;         // row, and the slot-sized group of rows it arrives with
        const int use = (int)__umulhi((uint32_t)g, inv_slots);
        r.slot = g - use * n_slots;
        r.R = R; r.sub = sub; y_out = y;
        r.q0 = s_raw + ((r.slot << cx.rps_shift) + (n & ((1 << cx.rps_shift) - 1))) * slot_bytes;
        // positions [R*P1w, (R+1)*P1w) replace those CAP earlier, last read by tile (pos - CAP + P1w) / 128
        const int last_reader = ((R + 2) * P1w - 1 - CAP) >> 7;      // arithmetic shift: negative = none
        if (last_reader >= tiles_waited) {
            mbar_wait(&tile_done[last_reader & (TILE_RING - 1)], (last_reader / TILE_RING) & 1);
            tiles_waited = last_reader + 1;
        }
        if (r.real) {
            int ld, want;                                // group g is loader g % n_loaders' (g / n_loaders + 1)-th
            if (n_loaders == 3) { ld = g % 3; want = g / 3 + 1; }
            else { ld = g & (n_loaders - 1); want = (g >> (n_loaders - 1)) + 1; }     // 1 or 2 loaders
            while (ld_acquire_shared(&s_rows_issued[ld]) < want) __nanosleep(32);
            mbar_wait(&raw_full[r.slot], use & 1);
        }
        sub += UNFOLD_WARPS % 3; R += UNFOLD_WARPS / 3; py += UNFOLD_WARPS / 3;
        if (sub >= 3) { sub -= 3; ++R; ++py; }
    };
    auto emit = [&](const Row &r, uint32_t (&w)[NP][5]) {
        const int pos0 = r.R * P1w + lane;
#pragma unroll
        for (int part = 0; part < NP; ++part) {
            const int px = part * 32 + lane;
            if (px < P1w) {
                uint4 lo, hi;
                chunk_from_raw<ACC16>(w[part], lo, hi);                     // zero rows: w = 0 gives the format's zeros
                const int pos = RG::wrap(pos0 + part * 32);
                uint8_t *dst = s_ring + r.sub * RG::SUB + pos * 16;
                *reinterpret_cast<uint4 *>(dst) = lo;
                *reinterpret_cast<uint4 *>(dst + RG::PLANE) = hi;
                if (pos < 128) {                                     // mirror past the end of the ring
                    *reinterpret_cast<uint4 *>(dst + CAP * 16) = lo;
                    *reinterpret_cast<uint4 *>(dst + CAP * 16 + RG::PLANE) = hi;
                }
            }
        }
    };
    auto gather = [&](const Row &r, uint32_t (&w)[NP][5]) {
        // (lanes past the last column and lane 0's tap -1 read a few bytes outside the row: still this CTA's shared memory)
#pragma unroll
        for (int part = 0; part < NP; ++part)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(r.q0 + part * PB + off0[j]);
                w[part][j] = r.real ? __byte_perm(wp[0], wp[1], sel[j]) : 0u;
            }
        if (lane == 0) w[0][0] = 0u;                                 // the padding column left of the image
    };
    // (taking two rows per step -- a second row's waits and loads in flight -- was tried: no gain, the SM's ALU pipe is the limit)
    for (int u = pwarp; u < total_u; u += UNFOLD_WARPS) {
        Row r0;
        int y0;
        const bool stamp = tl && pwarp == 0 && lane == 0 && rows_done < 96;
        if (stamp) tl[256 + 4 * rows_done] = clock64();
        next_row(r0, y0);
        if (stamp) tl[256 + 4 * rows_done + 1] = clock64();
        uint32_t w0[NP][5];
        if (GATHER) {
            gather(r0, w0);
        } else {
            // the resized row, each pixel once, into this warp's buffer; then five consecutive words per pooled column
            uint32_t *cmp = s_cmp + pwarp * F1_CMP_STRIDE;
            if (r0.real) {
                const uint8_t *q1 = r0.q0 + (src.n_src - 1) * src.row_bytes;
                const int b0 = s_yb[2 * y0], b1 = s_yb[2 * y0 + 1];
                if (plan.mode == RESIZE_LINEAR && plan.gather_step_x == 0) {
                    // Four pixels per lane at a time, results held in registers until all four are done: one pixel is a chain of
                    // ~12 dependent steps behind two rounds of shared-memory loads, and a store to cmp between two pixels would
                    // order the next pixel's loads behind it (q0 is a byte pointer: it may alias anything).  One pixel at a time
                    // with the dispatch on the plan inside the loop took ~3,600 cycles per 1080p row and warp.
                    const int last = plan.dst_w - 1;
                    for (int xb = lane; xb < plan.dst_w; xb += 32 * F1_RESIZE_ILP) {
                        uint32_t v[F1_RESIZE_ILP];
#pragma unroll
                        for (int k = 0; k < F1_RESIZE_ILP; ++k) v[k] = bilinear_word(r0.q0, q1, s_xtab, b0, b1, min(xb + 32 * k, last));
#pragma unroll
                        for (int k = 0; k < F1_RESIZE_ILP; ++k)
                            if (xb + 32 * k <= last) cmp[1 + xb + 32 * k] = v[k];
                    }
                } else {
                    for (int x = lane; x < plan.dst_w; x += 32) cmp[1 + x] = resized_word(plan, r0.q0, q1, s_xtab, b0, b1, x);
                }
            }
            __syncwarp();
#pragma unroll
            for (int part = 0; part < NP; ++part) {
                const int px = part * 32 + lane;
#pragma unroll
                for (int j = 0; j < 5; ++j) w0[part][j] = (r0.real && px < P1w) ? cmp[3 * px + j] : 0u;
            }
        }
        emit(r0, w0);
        // one fence for both directions: the stores above become visible to the MMA's async-proxy reads, and the loads from
        // the raw slot are ordered before the async-proxy refill that the release below allows
        if (stamp) tl[256 + 4 * rows_done + 2] = clock64();
        fence_proxy_async();
        __syncwarp();
        if (stamp) tl[256 + 4 * rows_done + 3] = clock64();
        ++rows_done;
        if (lane == 0) {
            if (r0.real) mbar_arrive(&raw_empty[r0.slot]);
            st_relaxed_shared(&s_rows_done[pwarp], rows_done);      // (every lane fenced its stores before the __syncwarp)
        }
    }
}

template <int C, bool GATHER, bool ACC16>
__global__ void __launch_bounds__(F1Roles<ACC16>::THREADS, 1) conv1_fused_tc_kernel(const Conv1Params p, const FusedSrc src) {
    using RL = F1Roles<ACC16>;
    constexpr int UNFOLD_WARPS = RL::UNFOLD_WARPS, MMA_WARPS = RL::MMA_WARPS, F1_MMA_WARP = RL::MMA_WARP0, F1_LOAD_WARP0 = RL::LOAD_WARP0;
    using S = F1Smem<C, UNFOLD_WARPS>;
    constexpr int CG = C / 8, CH = C / 2;
    constexpr int NP = (F_MAX_DST / 3 + 31) / 32;                               // 32-column parts of a resized row (P1w <= 85)
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *s_w = smem;
    uint8_t *s_ring = smem + S::OFF_RING;
    uint8_t *s_raw = smem + S::OFF_RAW;
    int *s_rowoff = reinterpret_cast<int *>(smem + S::OFF_TAB);               // [y][2]
    int *s_yb = s_rowoff + 2 * F_MAX_DST;                                       // [y][2]: b0, b1
    int4 *s_xtab = reinterpret_cast<int4 *>(s_yb + 2 * F_MAX_DST);              // [x]: 3*x0, 3*x1, a0, a1
    uint32_t *s_cmp = reinterpret_cast<uint32_t *>(smem + S::OFF_CMP);          // [unfold warp][F1_CMP_STRIDE]
    uint64_t *acc_full = reinterpret_cast<uint64_t *>(smem + S::OFF_BAR);
    uint64_t *acc_empty = acc_full + 3;
    uint64_t *raw_full = acc_empty + 3;                                         // [RAW_SLOTS_MAX]
    uint64_t *raw_empty = raw_full + RAW_SLOTS_MAX;                             // [RAW_SLOTS_MAX]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(raw_empty + RAW_SLOTS_MAX);
    int *s_rows_done = reinterpret_cast<int *>(tmem_slot + 2);                  // [UNFOLD_WARPS] rows finished by each unfold warp
    int *s_rows_issued = s_rows_done + UNFOLD_WARPS;                            // [LOADER_WARPS] rows issued by each loader
    uint64_t *tile_done = reinterpret_cast<uint64_t *>(s_rows_issued + 4);      // [TILE_RING] tile t's MMAs have completed
    float *s_par = reinterpret_cast<float *>(smem + S::OFF_PAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ResizePlanDev &plan = src.plan;
    const int H = p.H, P1w = p.P1w, RPF = p.P1h + 1;          // pooled rows per frame incl. the zero row
    const int Hc = min(H, 3 * p.P1h + 1);                     // resized rows the conv reads (row 3*P1h only if it exists)
    const int n_frames_cta = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_tiles = (n_frames_cta * RPF * P1w + 127) / 128;
    const int total_u = 3 * n_frames_cta * RPF;               // resized rows incl. the zero rows, u = 3R + sub
    const int n_slots = src.n_slots, slot_bytes = src.n_src * src.row_bytes;
    const uint32_t inv_slots = 0xffffffffu / (uint32_t)n_slots + 1u;       // n / n_slots = umulhi(n, inv_slots), exact while n * n_slots < 2^32

    for (int i = threadIdx.x; i < S::W_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(s_w)[i] = (ACC16 ? p.w_perm16 : p.w_perm)[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        if (ACC16) { reinterpret_cast<uint32_t *>(s_par)[i] = p.par16[i]; }
        else { s_par[i] = p.bias[i]; s_par[C + i] = p.scale_magic[i]; s_par[2 * C + i] = p.shift[i]; }
    }
    // The only positions read before they are written: the row above the first frame (tile 0's view shifted by -P1w).
    // two zero pixel values in the operand format; the last half of the second plane is the constant that multiplies the bias row
    const uint32_t Z2 = ACC16 ? 0u : 0x64006400u, Z2K = ACC16 ? 0x3c000000u : 0x64006400u;
    for (int i = threadIdx.x; i < UNFOLD_WARPS * F1_CMP_STRIDE; i += blockDim.x) s_cmp[i] = 0u;
    for (int i = threadIdx.x; i < 2 * P1w; i += blockDim.x)
        reinterpret_cast<uint4 *>(s_ring + 2 * FR_SUB + (i >= P1w ? FR_PLANE : 0))[FR_CAP - P1w + (i >= P1w ? i - P1w : i)] =
            make_uint4(Z2, Z2, Z2, i >= P1w ? Z2K : Z2);
    for (int y = threadIdx.x; y < H; y += blockDim.x) {
        int r0, r1, b0 = 2048, b1 = 0;
        if (plan.gather_step_x > 0) { r0 = r1 = plan.gather_off_y + y * plan.gather_step_y; }
        else if (plan.mode == RESIZE_COPY) { r0 = r1 = y; }
        else if (plan.mode == RESIZE_AREA2) { r0 = 2 * y; r1 = 2 * y + 1; }
        else { r0 = plan.y0[y]; r1 = plan.y1[y]; b0 = plan.b0[y]; b1 = plan.b1[y]; }
        if (src.compact) { r0 = plan.row_slot[r0]; r1 = (plan.mode == RESIZE_LINEAR && b1 == 0) ? r0 : plan.row_slot[r1]; }
        s_rowoff[2 * y] = (int)(r0 * src.row_pitch);
        s_rowoff[2 * y + 1] = (int)(r1 * src.row_pitch);
        s_yb[2 * y] = b0;
        s_yb[2 * y + 1] = b1;
    }
    if (plan.mode == RESIZE_LINEAR && plan.gather_step_x == 0)
        for (int x = threadIdx.x; x < plan.dst_w; x += blockDim.x) {
            // the second tap is the next pixel or (clamped at the edge) the same one: fold that case into the first coefficient, so
            // the kernel always reads the six bytes of pixels x0 and x0 + 1
            const int x0 = plan.x0[x], x1 = plan.x1[x];
            int a0 = plan.a0[x], a1 = plan.a1[x];
            if (x1 != x0 + 1) { a0 += a1; a1 = 0; }
            s_xtab[x] = make_int4(3 * x0, a0 | (a1 << 16), 0, 0);
        }
    if (threadIdx.x < UNFOLD_WARPS + LOADER_WARPS) s_rows_done[threadIdx.x] = 0;    // ... and s_rows_issued
    fence_proxy_async();
    if (threadIdx.x == 0) {
        for (int d = 0; d < 3; ++d) { mbar_init(&acc_full[d], 1); mbar_init(&acc_empty[d], EPI_WARPS); }
        for (int s = 0; s < RAW_SLOTS_MAX; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 1); }
        for (int s = 0; s < TILE_RING; ++s) mbar_init(&tile_done[s], MMA_WARPS);
        fence_barrier_init();
    }
    if (warp == F1_MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    long long *tl = (p.timeline && blockIdx.x == 0) ? p.timeline : nullptr;     // debug stamps of CTA 0 (cutdet_net_debug_timeline)
    if (tl && threadIdx.x == 0) tl[2047] = clock64();
    if (p.timeline && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); p.timeline[2048 + 2 * blockIdx.x] = g; }
    // Programmatic dependent launch (launch_conv1_fused): the kernel before this one -- conv2 of the previous sub-batch -- may
    // still be READING the activation buffer this kernel writes.  Loaders, unfold and MMAs do not touch it and start at once;
    // the epilogue warps wait for that kernel to complete before their first store.  The next kernel may be scheduled now.
    grid_dep_launch();

    if (warp < 8) reg_alloc<RL::REGS_EPI>();                     // one instruction per warpgroup (.sync.aligned)
    else if (warp >= RL::WARPS - 4) reg_dealloc<RL::REGS_LIGHT>();
    else reg_dealloc<RL::REGS_UNFOLD>();
    if (warp >= F1_MMA_WARP) {
        if (warp >= F1_LOAD_WARP0) {
            // ------------------------------------------------------------------ loaders: source rows -> raw ring
            // (the wait on raw_empty is for row n - n_slots to have been read; the slot cannot be a phase further, since that
            // takes row n itself)
            const int lw = warp - F1_LOAD_WARP0, total_rows = n_frames_cta * Hc;
            const int n_loaders = min(LOADER_WARPS, n_slots);
            if (lw < n_loaders) {
                const uint64_t stream_once = l2_policy_evict_first();   // frames are read once: keep the L2 for the activations
                int fi = 0, y = lw, issued = 0;
                for (int n = lw; n < total_rows; n += n_loaders, y += n_loaders) {
                    while (y >= Hc) { y -= Hc; ++fi; }
                    const int use = (int)__umulhi((uint32_t)n, inv_slots), slot = n - use * n_slots;
                    mbar_wait(&raw_empty[slot], (use & 1) ^ 1);
                    if (tl && lane == 0 && n < 256) tl[n] = clock64();
                    ++issued;
                    if (elect_one()) {
                        const uint8_t *frame = src.frames + (long long)(blockIdx.x + (long long)fi * gridDim.x) * src.frame_stride;
                        mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)slot_bytes);
                        for (int j = 0; j < src.n_src; ++j)
                            bulk_load_1d_hint(s_raw + slot * slot_bytes + j * src.row_bytes, frame + s_rowoff[2 * y + j],
                                              (uint32_t)src.row_bytes, &raw_full[slot], stream_once);
                        st_release_shared(&s_rows_issued[lw], issued);   // raw_full[slot] is now in row n's phase
                    }
                    __syncwarp();
                }
            }
        } else {
            // ------------------------------------------------------------------ MMA issuers: warp m takes block rows dy % MMA_WARPS == m
            const int m_warp = warp - F1_MMA_WARP;
            const bool tl_mma = tl && m_warp == 0 && lane == 0;
            uint32_t acc_phase = 0;
            const uint32_t w_addr = smem_u32(s_w), ring_addr = smem_u32(s_ring);
            const uint32_t idesc = ACC16 ? instr_desc_f16_acc16(128, 3 * C) : instr_desc_16bit(128, 3 * C, kBf16);
            int r_hi = 127 / P1w, r_rem = 127 % P1w;      // pooled row of the tile's last position, kept without a division per tile
            for (int t = 0; t < n_tiles; ++t) {
                // rows of sub-rings 1 and 2 up to pooled row (128t + 127) / P1w, of sub-ring 0 one pooled row further
                const int u_hi = min(total_u - 1, 3 * (r_hi + 1));
                r_rem += 128;                                 // 64 <= P1w: at most three rows further
    #pragma unroll
                for (int k = 0; k < 3; ++k) { const bool c = r_rem >= P1w; r_rem -= c ? P1w : 0; r_hi += c ? 1 : 0; }
                const int mine = lane % UNFOLD_WARPS;
                const int need = u_hi >= mine ? (u_hi - mine) / UNFOLD_WARPS + 1 : 0;
                if (tl_mma && t < 60) tl[1800 + 4 * t + 3] = clock64();
                while (!__all_sync(0xffffffffu, ld_acquire_shared(&s_rows_done[mine]) >= need)) __nanosleep(32);
                __syncwarp();
                tc_fence_after_sync();
                if (tl_mma && t < 64) tl[1024 + t] = clock64();
                uint32_t a_chunk[5];
    #pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const int sub = (c + 2) % 3, shift = c == 0 ? -P1w : (c == 4 ? P1w : 0);
                    a_chunk[c] = ring_addr + sub * FR_SUB + (uint32_t)((t * 128 + shift) & (FR_CAP - 1)) * 16;
                }
    #pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    if (MMA_WARPS > 1 && dy != m_warp) continue;
                    mbar_wait(&acc_empty[dy], acc_phase ^ 1);
                    tc_fence_after_sync();
                    if (tl_mma && t < 60) tl[1800 + 4 * t + dy] = clock64();
                    if (elect_one()) {
    #pragma unroll
                        for (int ks = 0; ks < 3; ++ks) {
                            const uint64_t da = smem_desc(a_chunk[dy + ks], FR_PLANE, 128);
                            const uint64_t db = smem_desc(w_addr + 2 * ks * S::LBO_B, S::LBO_B, 128);
                            umma_16bit(tmem_base + 3 * C * dy, da, db, idesc, ks > 0 ? 1u : 0u);
                        }
                        umma_commit(&acc_full[dy]);
                        // tile t no longer reads the operand ring once every issuer's MMAs of it have completed
                        if (MMA_WARPS > 1 || dy == 2) umma_commit(&tile_done[t & (TILE_RING - 1)]);
                    }
                    __syncwarp();
                }
                if (tl_mma && t < 64) tl[1088 + t] = clock64();
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 8 && warp < 8 + UNFOLD_WARPS) {
        // ------------------------------------------------------------------ unfold: raw rows -> x-unfolded fp16 ring
        F1Ctx cx{&p, &src, s_ring, s_raw, s_cmp, s_yb, s_xtab, raw_full, raw_empty, tile_done, s_rows_done, s_rows_issued,
                 RPF, Hc, total_u, n_slots, slot_bytes, min(LOADER_WARPS, n_slots), inv_slots, tl};
        f1_unfold_role<C, GATHER, ACC16, UNFOLD_WARPS>(cx, warp - 8, lane);
    } else {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3, half = warp >> 2, m = q * 32 + lane, ch0 = half * CH;
        const uint32_t tmem_thread = tmem_base + ((uint32_t)(q * 32) << 16) + half * 3 * CH;   // columns [dy][half][dx][CH]
        uint32_t acc_phase = 0;
        grid_dep_wait();
        for (int fi = 0; fi < n_frames_cta; ++fi) zero_pads(p.out, CG, blockIdx.x + fi * gridDim.x, blockIdx.x + fi * gridDim.x + 1, threadIdx.x, EPI_WARPS * 32);
        int X = m % P1w, Y = m / P1w, fi = 0;              // position 128 t + m = ((fi * RPF + Y) * P1w + X), advanced tile by tile
        while (Y >= RPF) { Y -= RPF; ++fi; }
        using Row = typename std::conditional<ACC16, EpiRow16<CH>, EpiRow<CH>>::type;
        Row bufA, bufB;
        if (n_tiles > 0) {
            mbar_wait(&acc_full[0], 0);
            tc_fence_after_sync();
            if constexpr (ACC16) epi_issue_row16<C>(tmem_thread, 0, bufA);
            else epi_issue_row<C>(tmem_thread, 0, bufA);
        }
        auto advance = [&]() {
            // next tile: 128 positions on.  64 <= P1w (fused_source), so at most three rows further; branch-free, the lanes differ
            X += 128;
#pragma unroll
            for (int k = 0; k < 3; ++k) { const bool c = X >= P1w; X -= c ? P1w : 0; Y += c ? 1 : 0; }
#pragma unroll
            for (int k = 0; k < 3; ++k) { const bool c = Y >= RPF; Y -= c ? RPF : 0; fi += c ? 1 : 0; }
        };
        auto tile = [&](int t, Row &cur, Row &nxt) {
            const bool valid = fi < n_frames_cta && Y < p.P1h;
            if constexpr (ACC16) {
                // (a table of store offsets by position inside the frame, fetched a tile ahead, was tried instead of this
                // div/mod chain: no gain)
                uint4 *dst = store_addr16<C>(p.out, blockIdx.x + fi * gridDim.x, Y, X, ch0);
                advance();
                uint32_t v[CH / 2];
                bool requested = epilogue_tile_pipelined16<C>(tmem_thread, acc_full, acc_empty, acc_phase, lane, t + 1 < n_tiles, cur, nxt, v,
                                                              (tl && threadIdx.x == 0 && t < 64) ? tl + 1216 + 8 * t : nullptr);
                epilogue_scale_shift16<C>(reinterpret_cast<const uint32_t *>(s_par), ch0, v);
                if (!requested && mbar_test_wait(&acc_full[0], acc_phase ^ 1)) {      // second chance: the load flies behind the store
                    tc_fence_after_sync();
                    epi_issue_row16<C>(tmem_thread, 0, nxt);
                    requested = true;
                }
                if (valid) store_pixel16<C>(dst, p.out.gtot, v);
                if (!requested) {
                    mbar_wait(&acc_full[0], acc_phase ^ 1);
                    tc_fence_after_sync();
                    epi_issue_row16<C>(tmem_thread, 0, nxt);
                }
            } else {
                float v[CH];
                epilogue_tile_pipelined<C>(tmem_thread, acc_full, acc_empty, acc_phase, lane, t + 1 < n_tiles, cur, nxt, v);
                epilogue_scale_shift<C>(s_par, ch0, v);
                if (valid) store_pixel<C>(p.out, blockIdx.x + fi * gridDim.x, Y, X, ch0, v);
                advance();
            }
            if (tl && threadIdx.x == 0 && t < 64) tl[1152 + t] = clock64();
            acc_phase ^= 1;
        };
        for (int t = 0; t < n_tiles; t += 2) {          // the two buffers swap roles from one tile to the next
            tile(t, bufA, bufB);
            if (t + 1 < n_tiles) tile(t + 1, bufB, bufA);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (p.timeline && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); p.timeline[2049 + 2 * blockIdx.x] = g; }
    if (warp == F1_MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ helpers of the two-set kernel
template <int C>
__device__ __forceinline__ void row_load_all(uint32_t taddr, uint32_t (&row)[3 * C / 2]) {
    static_assert(C == 48 || C == 32, "channels");
    if constexpr (C == 48) {            // 144 columns: 64 + 64 + 16 (an x64 load alone does not fit the kernel's launch-time register count)
        tmem_ld_pack32(taddr, row);
        tmem_ld_pack32(taddr + 64, row + 32);
        tmem_ld_pack8(taddr + 128, row + 64);
    } else {                            // 96 columns: 64 + 32
        tmem_ld_pack32(taddr, row);
        tmem_ld_pack16(taddr + 64, row + 32);
    }
}
// the row's columns are ordered [channel half][dx][C/2 channels]: max over dx, channel pairs in channel order
template <int C>
__device__ __forceinline__ void row_max_dx(const uint32_t (&row)[3 * C / 2], uint32_t (&m)[C / 2]) {
    constexpr int Q = C / 4;            // packed pairs per (half, dx)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < Q; ++j) m[h * Q + j] = hmax3(row[h * 3 * Q + j], row[h * 3 * Q + Q + j], row[h * 3 * Q + 2 * Q + j]);
}
// ------------------------------------------------------------------------------------------------ K1 + layer 1, two epilogue sets
// The fused kernel with TWO sets of epilogue warps that alternate over the tiles and an MMA issuer per block row.
// TMEM holds one accumulator tile (432 of 512 columns), so tile t + 1's block row dy can only be computed once tile t's row dy has
// been read -- but nothing requires the SAME warps to read consecutive tiles.  conv1_fused_tc_kernel's epilogue chain per tile
// (three rounds of barrier / tcgen05.ld / wait::ld / hand-back, then max, affine, address, store: ~1,450 cycles by its clock
// stamps) sets its tile period; with two sets that chain has two periods to complete, and what bounds a period is the loop of
// one block row: MMAs (216 cycles) -> commit -> a set reads the row (~350) -> hand-back -> issue.
// A set is FOUR warps (one per TMEM lane quarter; a thread takes all C channels of its pooled pixel: 3C/2 packed registers per
// block row): with 16 epilogue warps of half the channels each (1,024 threads) the epilogue was as quick, but the unfold warps,
// left with 64 registers among 32 warps, took 1,600 cycles per row and the rows became the bottleneck at the same 1,450 cycles.
//   warps 0-3 set A | 4-7 set B | 8-15 unfold | 16-18 MMA issuers (block row dy), 19 idle | 20-22 loaders, 23 idle
struct S2Roles {
    static constexpr int EPI_WARPS2 = 8, UNFOLD_WARPS = 8, UNFOLD_WARP0 = 8, MMA_WARP0 = 16, LOAD_WARP0 = 20, WARPS = 24, THREADS = 32 * WARPS;
    static constexpr int REGS_START = 80, REGS_EPI = 120, REGS_UNFOLD = 88, REGS_MMA = 32, REGS_LOAD = 32;
    static_assert(THREADS * REGS_START == 256 * REGS_EPI + 256 * REGS_UNFOLD + 128 * REGS_MMA + 128 * REGS_LOAD, "register pool");
};

template <int C, bool GATHER>
__global__ void __launch_bounds__(S2Roles::THREADS, 1) conv1_fused_sets_kernel(const Conv1Params p, const FusedSrc src) {
    using RL = S2Roles;
    constexpr int UNFOLD_WARPS = RL::UNFOLD_WARPS;
    using S = F1Smem<C, UNFOLD_WARPS>;
    constexpr int CG = C / 8, NPK = C / 2, ROWPK = 3 * C / 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *s_w = smem;
    uint8_t *s_ring = smem + S::OFF_RING;
    uint8_t *s_raw = smem + S::OFF_RAW;
    int *s_rowoff = reinterpret_cast<int *>(smem + S::OFF_TAB);               // [y][2]
    int *s_yb = s_rowoff + 2 * F_MAX_DST;                                       // [y][2]: b0, b1
    int4 *s_xtab = reinterpret_cast<int4 *>(s_yb + 2 * F_MAX_DST);              // [x]: 3*x0, a0 | a1 << 16
    uint32_t *s_cmp = reinterpret_cast<uint32_t *>(smem + S::OFF_CMP);          // [unfold warp][F1_CMP_STRIDE]
    uint64_t *acc_full = reinterpret_cast<uint64_t *>(smem + S::OFF_BAR);       // [set][3]: block row dy of the set's tile is complete
    uint64_t *acc_rel = acc_full + 6;                                           // [set][3]: ... has been read by the set (4 warps)
    uint64_t *raw_full = acc_rel + 6;                                           // [RAW_SLOTS_MAX]
    uint64_t *raw_empty = raw_full + RAW_SLOTS_MAX;                             // [RAW_SLOTS_MAX]
    uint64_t *tile_done = raw_empty + RAW_SLOTS_MAX;                            // [TILE_RING] tile t's MMAs have completed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tile_done + TILE_RING);
    int *s_rows_done = reinterpret_cast<int *>(tmem_slot + 2);                  // [UNFOLD_WARPS] rows finished by each unfold warp
    int *s_rows_issued = s_rows_done + UNFOLD_WARPS;                            // [LOADER_WARPS] rows issued by each loader
    static_assert((6 + 6 + 2 * RAW_SLOTS_MAX + TILE_RING) * 8 + 8 + (UNFOLD_WARPS + 4) * 4 <= 1024, "barrier block");
    uint32_t *s_par16 = reinterpret_cast<uint32_t *>(smem + S::OFF_PAR);        // [C/2 scale pairs | C/2 shift pairs]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ResizePlanDev &plan = src.plan;
    const int H = p.H, P1w = p.P1w, RPF = p.P1h + 1;          // pooled rows per frame incl. the zero row
    const int Hc = min(H, 3 * p.P1h + 1);                     // resized rows the conv reads (row 3*P1h only if it exists)

    for (int i = threadIdx.x; i < S::W_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(s_w)[i] = p.w_perm16[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_par16[i] = p.par16[i];
    // operand-format zeros for the row above the first frame (tile 0's view shifted by -P1w); the last half of the second
    // plane is the constant 1.0 that multiplies the bias row
    for (int i = threadIdx.x; i < UNFOLD_WARPS * F1_CMP_STRIDE; i += blockDim.x) s_cmp[i] = 0u;
    for (int i = threadIdx.x; i < 2 * P1w; i += blockDim.x)
        reinterpret_cast<uint4 *>(s_ring + 2 * FR_SUB + (i >= P1w ? FR_PLANE : 0))[FR_CAP - P1w + (i >= P1w ? i - P1w : i)] =
            make_uint4(0u, 0u, 0u, i >= P1w ? 0x3c000000u : 0u);
    for (int y = threadIdx.x; y < H; y += blockDim.x) {
        int r0, r1, b0 = 2048, b1 = 0;
        if (plan.gather_step_x > 0) { r0 = r1 = plan.gather_off_y + y * plan.gather_step_y; }
        else if (plan.mode == RESIZE_COPY) { r0 = r1 = y; }
        else if (plan.mode == RESIZE_AREA2) { r0 = 2 * y; r1 = 2 * y + 1; }
        else { r0 = plan.y0[y]; r1 = plan.y1[y]; b0 = plan.b0[y]; b1 = plan.b1[y]; }
        if (src.compact) { r0 = plan.row_slot[r0]; r1 = (plan.mode == RESIZE_LINEAR && b1 == 0) ? r0 : plan.row_slot[r1]; }
        s_rowoff[2 * y] = (int)(r0 * src.row_pitch);
        s_rowoff[2 * y + 1] = (int)(r1 * src.row_pitch);
        s_yb[2 * y] = b0;
        s_yb[2 * y + 1] = b1;
    }
    if (plan.mode == RESIZE_LINEAR && plan.gather_step_x == 0)
        for (int x = threadIdx.x; x < plan.dst_w; x += blockDim.x) {
            const int x0 = plan.x0[x], x1 = plan.x1[x];
            int a0 = plan.a0[x], a1 = plan.a1[x];
            if (x1 != x0 + 1) { a0 += a1; a1 = 0; }          // clamped at the edge: always read the six bytes of pixels x0, x0 + 1
            s_xtab[x] = make_int4(3 * x0, a0 | (a1 << 16), 0, 0);
        }
    if (threadIdx.x < UNFOLD_WARPS + LOADER_WARPS) s_rows_done[threadIdx.x] = 0;       // ... and s_rows_issued
    fence_proxy_async();
    if (threadIdx.x == 0) {
        for (int d = 0; d < 6; ++d) { mbar_init(&acc_full[d], 1); mbar_init(&acc_rel[d], 4); }
        for (int s = 0; s < RAW_SLOTS_MAX; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 1); }
        for (int s = 0; s < TILE_RING; ++s) mbar_init(&tile_done[s], 3);
        fence_barrier_init();
    }
    if (warp == RL::MMA_WARP0) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // (declared after the set-up on purpose: the kernel starts with 64 registers per thread, and values that live across the
    // set-up were spilled there and re-read from L2 inside the roles' loops)
    const int n_frames_cta = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_tiles = (n_frames_cta * RPF * P1w + 127) / 128;
    const int total_u = 3 * n_frames_cta * RPF;               // resized rows incl. the zero rows, u = 3R + sub
    const int n_slots = src.n_slots, slot_bytes = src.n_src * src.row_bytes;
    const uint32_t inv_slots = 0xffffffffu / (uint32_t)n_slots + 1u;       // n / n_slots = umulhi(n, inv_slots), exact while n * n_slots < 2^32
    long long *tl = (p.timeline && blockIdx.x == 0) ? p.timeline : nullptr;     // debug stamps of CTA 0 (cutdet_net_debug_timeline)
    if (tl && threadIdx.x == 0) tl[2047] = clock64();
    if (p.timeline && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); p.timeline[2048 + 2 * blockIdx.x] = g; }
    // Programmatic dependent launch: the kernel before this one -- conv2 of the previous sub-batch -- may still be READING the
    // activation buffer this kernel writes.  Only the epilogue's stores and the pad zeroing touch it: they wait for that kernel
    // to complete (grid_dep_wait), everything else starts at once.  The next kernel may be scheduled now.
    grid_dep_launch();

    // Register budgets per warpgroup (setmaxnreg.sync.aligned), set at the top of each role's branch so that ptxas sees which
    // budget governs which code (set in a separate if-chain it allocated every role within the smallest one).
    if (warp >= RL::LOAD_WARP0) {
      reg_dealloc<RL::REGS_LOAD>();
      if (warp < RL::LOAD_WARP0 + LOADER_WARPS) {
        // ------------------------------------------------------------------ loaders: source rows -> raw ring
        // row n is issued by loader n % n_loaders into slot n % n_slots (the slot count is a multiple of the loader count, so a
        // slot is always refilled by the same loader; the wait on raw_empty is for row n - n_slots to have been read)
        const int lw = warp - RL::LOAD_WARP0, total_rows = n_frames_cta * Hc;
        const int n_loaders = min(LOADER_WARPS, n_slots);
        if (lw < n_loaders) {
            const uint64_t stream_once = l2_policy_evict_first();   // frames are read once: keep the L2 for the activations
            int fi = 0, y = lw, issued = 0;
            for (int n = lw; n < total_rows; n += n_loaders, y += n_loaders) {
                while (y >= Hc) { y -= Hc; ++fi; }
                const int use = (int)__umulhi((uint32_t)n, inv_slots), slot = n - use * n_slots;
                mbar_wait(&raw_empty[slot], (use & 1) ^ 1);
                if (tl && lane == 0 && n < 256) tl[n] = clock64();
                ++issued;
                if (elect_one()) {
                    const uint8_t *frame = src.frames + (long long)(blockIdx.x + (long long)fi * gridDim.x) * src.frame_stride;
                    mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)slot_bytes);
                    for (int j = 0; j < src.n_src; ++j)
                        bulk_load_1d_hint(s_raw + slot * slot_bytes + j * src.row_bytes, frame + s_rowoff[2 * y + j],
                                          (uint32_t)src.row_bytes, &raw_full[slot], stream_once);
                    st_release_shared(&s_rows_issued[lw], issued);   // raw_full[slot] is now in row n's phase
                }
                __syncwarp();
            }
        }
      }
    } else if (warp >= RL::MMA_WARP0) {
      reg_dealloc<RL::REGS_MMA>();
      if (warp < RL::MMA_WARP0 + 3) {
        // ------------------------------------------------------------------ MMA issuer of block row dy
        const int dy = warp - RL::MMA_WARP0;
        const bool stamp = tl && lane == 0;
        const uint32_t w_addr = smem_u32(s_w), ring_addr = smem_u32(s_ring);
        const uint32_t idesc = instr_desc_f16_acc16(128, 3 * C);
        const uint32_t tmem_row = tmem_base + 3 * C * dy;
        int r_hi = 127 / P1w, r_rem = 127 % P1w;      // pooled row of the tile's last position, kept without a division per tile
        for (int t = 0; t < n_tiles; ++t) {
            // block row dy reads input rows 3R - 1 + dy .. 3R + 1 + dy of pooled row R: resized rows up to u = 3 r_hi + 1 + dy
            const int u_hi = min(total_u - 1, 3 * r_hi + 1 + dy);
            r_rem += 128;                                 // 64 <= P1w: at most three rows further
#pragma unroll
            for (int k = 0; k < 3; ++k) { const bool c = r_rem >= P1w; r_rem -= c ? P1w : 0; r_hi += c ? 1 : 0; }
            const int mine = lane % UNFOLD_WARPS;
            const int need = u_hi >= mine ? (u_hi - mine) / UNFOLD_WARPS + 1 : 0;
            while (!__all_sync(0xffffffffu, ld_acquire_shared(&s_rows_done[mine]) >= need)) __nanosleep(20);
            __syncwarp();
            if (stamp && dy == 0 && t < 64) tl[1024 + t] = clock64();
            uint32_t a_chunk[3];
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                const int c = dy + ks, sub = (c + 2) % 3, shift = c == 0 ? -P1w : (c == 4 ? P1w : 0);
                a_chunk[ks] = ring_addr + sub * FR_SUB + (uint32_t)((t * 128 + shift) & (FR_CAP - 1)) * 16;
            }
            // block row dy of the accumulator was last filled for tile t - 1, which the OTHER set reads: its (t - 1) / 2-th tile
            if (t > 0) mbar_wait(&acc_rel[3 * ((t - 1) & 1) + dy], (uint32_t)((t - 1) >> 1) & 1u);
            tc_fence_after_sync();
            if (stamp && t < 60) tl[1800 + 4 * t + dy] = clock64();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) {
                    const uint64_t da = smem_desc(a_chunk[ks], FR_PLANE, 128);
                    const uint64_t db = smem_desc(w_addr + 2 * ks * S::LBO_B, S::LBO_B, 128);
                    umma_16bit(tmem_row, da, db, idesc, ks > 0 ? 1u : 0u);
                }
                umma_commit(&acc_full[3 * (t & 1) + dy]);
                umma_commit(&tile_done[t & (TILE_RING - 1)]);     // tile t no longer reads the operand ring once all three rows have completed
            }
            __syncwarp();
        }
      }
    } else if (warp >= RL::UNFOLD_WARP0) {
        // ------------------------------------------------------------------ unfold: raw rows -> x-unfolded fp16 ring
        if constexpr (RL::REGS_UNFOLD > RL::REGS_START) reg_alloc<RL::REGS_UNFOLD>();      // (waits for what the light groups give up)
        else reg_dealloc<RL::REGS_UNFOLD>();
        F1Ctx cx{&p, &src, s_ring, s_raw, s_cmp, s_yb, s_xtab, raw_full, raw_empty, tile_done, s_rows_done, s_rows_issued,
                 RPF, Hc, total_u, n_slots, slot_bytes, min(LOADER_WARPS, n_slots), inv_slots, tl};
        f1_unfold_role<C, GATHER, true, UNFOLD_WARPS>(cx, warp - RL::UNFOLD_WARP0, lane);
        // ... and, once the rows are through, the entries of the output buffer that are not pixels (they belong to no GEMM row)
        grid_dep_wait();
        for (int fi = 0; fi < n_frames_cta; ++fi)
            zero_pads(p.out, CG, blockIdx.x + fi * gridDim.x, blockIdx.x + fi * gridDim.x + 1, threadIdx.x - 32 * RL::UNFOLD_WARP0, 32 * UNFOLD_WARPS);
    } else {
        // ------------------------------------------------------------------ epilogue sets: set s takes tiles s, s + 2, s + 4, ...
        // A thread is a TMEM lane (a pooled pixel) and ALL its channels.  A set's chain per tile -- three rounds of barrier, TMEM
        // load, wait::ld and hand-back, then ReLU/affine, address and store -- has two tile periods to complete, and block row dy
        // of the next tile (the other set's) is issued the moment this set has handed row dy back.
        reg_alloc<RL::REGS_EPI>();
        const int set = warp >> 2, q = warp & 3, mrow = q * 32 + lane;
        const uint32_t tm_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint64_t *full = &acc_full[3 * set], *rel = &acc_rel[3 * set];
        // (no debug stamps here: the set has no register to spare, and a spilled word costs an L2 round trip per tile)
        int X = (128 * set + mrow) % p.P1w, Y = (128 * set + mrow) / p.P1w, fi = 0;   // position 128 t + mrow = ((fi * RPF + Y) * P1w + X)
        while (Y >= p.P1h + 1) { Y -= p.P1h + 1; ++fi; }
        const uint4 *sc4 = reinterpret_cast<const uint4 *>(s_par16), *sh4 = reinterpret_cast<const uint4 *>(s_par16 + NPK);
        uint32_t par = 0;
        for (int t = set; t < n_tiles; t += 2, par ^= 1) {
            uint32_t run[NPK];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                mbar_wait(&full[dy], par);
                tc_fence_after_sync();
                uint32_t row[ROWPK], m[NPK];
                row_load_all<C>(tm_lane + 3 * C * dy, row);
                tmem_ld_wait();
                reg_fence_u<ROWPK>(row);
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&rel[dy]);                 // the next tile's MMAs of this block row may start
                row_max_dx<C>(row, m);
#pragma unroll
                for (int i = 0; i < NPK; ++i) run[i] = dy == 0 ? m[i] : (dy == 1 ? hmax2(run[i], m[i]) : hmax3(run[i], m[i], 0u));   // ... and the ReLU
            }
#pragma unroll
            for (int i = 0; i < NPK / 4; ++i) {                 // scale (+-2^k, exact) and BatchNorm shift
                const uint4 sc = sc4[i], sh = sh4[i];
                run[4 * i + 0] = hfma2(run[4 * i + 0], sc.x, sh.x);
                run[4 * i + 1] = hfma2(run[4 * i + 1], sc.y, sh.y);
                run[4 * i + 2] = hfma2(run[4 * i + 2], sc.z, sh.z);
                run[4 * i + 3] = hfma2(run[4 * i + 3], sc.w, sh.w);
            }
            if (t == set) grid_dep_wait();                         // the kernel that may still read this buffer has completed
            if (fi < n_frames_cta && Y < p.P1h) {
                // entry of the phase-split buffer (32-bit arithmetic: the buffer has fewer than 2^32 16-byte entries)
                const uint32_t off = (uint32_t)(((Y % 3) * 3 + X % 3) * CG) * (uint32_t)p.out.gtot +
                                     (uint32_t)(p.out.frame0 + blockIdx.x + fi * gridDim.x) * (uint32_t)p.out.FP + (uint32_t)((Y / 3) * p.out.PW + X / 3);
                uint4 *dst = reinterpret_cast<uint4 *>(p.out.ptr) + off;
#pragma unroll
                for (int j = 0; j < CG; ++j) dst[(size_t)j * p.out.gtot] = make_uint4(run[4 * j], run[4 * j + 1], run[4 * j + 2], run[4 * j + 3]);
            }
            X += 256;                                              // this set's next tile: 64 <= P1w, at most four rows further; branch-free
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) { const bool c = X >= p.P1w; X -= c ? p.P1w : 0; Y += c ? 1 : 0; }
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) { const bool c = Y > p.P1h; Y -= c ? p.P1h + 1 : 0; fi += c ? 1 : 0; }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (p.timeline && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); p.timeline[2049 + 2 * blockIdx.x] = g; }
    if (warp == RL::MMA_WARP0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ K1 + layer 1 + layer 2, frame by frame
// One persistent CTA per SM takes a frame through BOTH layers before it starts the next one: the fused K1 + layer-1 pipeline of
// conv1_fused_tc_kernel (fp16 accumulators) writes the frame's pooled map into this CTA's slot of the phase-split buffer, then the
// same CTA runs layer 2's tiles of that slot (conv_mid_tc_kernel's pipeline: the first loader becomes the TMA producer, the issuer
// and the two other loaders issue one block row each, the eight epilogue warps serve both layers) and stores the frame's layer-2
// map for conv3.
// Why: as two kernels per 148-frame sub-batch the layers are joined by GRID-wide dependencies -- conv2 waits for the last conv1
// CTA, conv1's epilogue of the next sub-batch waits for the last conv2 CTA (it overwrites what conv2 reads) -- and every launch
// pays its ramp and drain: by the per-CTA clock stamps a sub-batch takes ~62 us of which ~46 us are the two tile loops.  A frame's
// layer-2 rows depend on that frame's layer-1 map only (the conv padding is per frame, and with the frame pitch a multiple of 128
// positions a frame is a whole number of layer-2 tiles), so nothing has to cross CTAs: here a phase change is a __syncthreads.
// The slot (400 KB at 720p) is written and read back by the same SM within ~50 us; the 148 slots (59 MB) stay in the L2.
// Shared memory is time-shared by the phases: layer 2's taps and stages overlay layer 1's taps and rings; the resize tables, all
// mbarriers and both layers' epilogue parameters live beyond layer 2's footprint and persist.  Layer 1 starts every frame from
// zero (its counters and barrier phases are re-initialised: the pipeline is quiescent at a phase change); layer 2's barrier
// phases simply run on from frame to frame.
// What was tried on layer 1 of conv12_frames and measured on one box (4,050 frames 720p, ms per call; all variants bit-equal;
// profiles/README.md, round 2), before and after the loaders moved to TMA boxes of two rows:
//   one issuer, eight epilogue warps on every tile, three loaders, one row per bulk copy ............ 1.449 -> 1.418 with TMA rows
//   the third loader issues block row 2 instead (two loaders) ...................................... 1.581
//   two epilogue SETS of four warps alternating over the tiles (as conv1_fused_sets_kernel) ........ 1.457 -> 1.427 with TMA rows
//   an issuer per block row (two of the eight unfold warps issue) .................................. 1.462 -> 1.506 with TMA rows
//   both ............................................................................................ 1.475 -> 1.508 with TMA rows
//   two / three row streams per loader warp (lanes issuing their own rows) ......................... 1.522 / 1.521 (base 1.556 in that build)
//   two / four rows per release of a loader's issue counter ........................................ 1.484 / 1.494 (base 1.457)
// Only the row supply moved the number: a thread gets a bulk copy accepted every ~500 cycles, so three loaders with one row per copy
// delivered 4.5 rows per ~1,400 cycles, exactly the tile period; with that lifted the stages are balanced again (fewer unfold
// warps or fewer loaders cost time, more issuers or epilogue sets gain none).
template <int C>
struct F12Smem {
    using S1 = F1Smem<C, F1Roles<true>::UNFOLD_WARPS>;
    using S2 = MidSmem<C>;
    static constexpr int OFF_BAR2 = S1::OFF_BAR + 640;             // layer 2's 14 barriers, inside layer 1's 1 KB barrier block
    static constexpr int OFF_PAR2 = S1::total;                     // layer 2's bias / scale / shift
    static constexpr int total = OFF_PAR2 + 3 * C * 4;
    static_assert(S2::W_BYTES + MID_STAGES * MID_STAGE_BYTES <= S1::OFF_TAB, "layer 2's operands must end before the tables layer 1 keeps");
    static_assert(2 * (6 + 2 * RAW_SLOTS_MAX + TILE_RING) * 4 + 8 + (F1Roles<true>::UNFOLD_WARPS + 4) * 4 <= 640, "layer 1's barrier block");
    static_assert(total <= 232448, "227 KB of shared memory per CTA");
};

template <int C, bool GATHER, int CAP = FR_CAP>
__global__ void __launch_bounds__(F1Roles<true>::THREADS, 1)
conv12_frames_kernel(const Conv1Params p, const FusedSrc src, const __grid_constant__ CUtensorMap in_map, const MidParams p2,
                     const __grid_constant__ CUtensorMap src_map) {
    using RL = F1Roles<true>;
    using RG = FrRing<CAP>;              // layer 1's operand ring; what it does not use of the FR_CAP layout belongs to the raw ring
    constexpr int UNFOLD_WARPS = RL::UNFOLD_WARPS, F1_MMA_WARP = RL::MMA_WARP0, F1_LOAD_WARP0 = RL::LOAD_WARP0;
    static_assert(RL::MMA_WARPS == 1, "one issuer warp in layer 1 (layer 2 adds two of the loaders)");
    using S = F1Smem<C, UNFOLD_WARPS>;
    using S2 = MidSmem<C>;
    using SS = F12Smem<C>;
    constexpr int CG = C / 8, CH = C / 2, KS = C / 16;
    extern __shared__ __align__(1024) uint8_t smem[];
    // ---- layer 1 (the layout of conv1_fused_tc_kernel)
    uint8_t *s_w = smem;
    uint8_t *s_ring = smem + S::OFF_RING;
    uint8_t *s_raw = smem + S::OFF_RAW - RG::EXTRA_RAW;
    int *s_rowoff = reinterpret_cast<int *>(smem + S::OFF_TAB);
    int *s_yb = s_rowoff + 2 * F_MAX_DST;
    int4 *s_xtab = reinterpret_cast<int4 *>(s_yb + 2 * F_MAX_DST);
    uint32_t *s_cmp = reinterpret_cast<uint32_t *>(smem + S::OFF_CMP);
    uint64_t *acc_full = reinterpret_cast<uint64_t *>(smem + S::OFF_BAR);
    uint64_t *acc_empty = acc_full + 3;
    uint64_t *raw_full = acc_empty + 3;
    uint64_t *raw_empty = raw_full + RAW_SLOTS_MAX;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(raw_empty + RAW_SLOTS_MAX);
    int *s_rows_done = reinterpret_cast<int *>(tmem_slot + 2);
    int *s_rows_issued = s_rows_done + UNFOLD_WARPS;                           // [LOADER_WARPS] slots issued by each loader
    uint64_t *tile_done = reinterpret_cast<uint64_t *>(s_rows_issued + 4);
    uint32_t *s_par = reinterpret_cast<uint32_t *>(smem + S::OFF_PAR);
    // ---- layer 2 (conv_mid_tc_kernel's operands over the same bytes; its barriers and parameters in the persistent tail)
    uint8_t *s_w2 = smem;
    uint8_t *s_stage = smem + S2::W_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + SS::OFF_BAR2);
    uint64_t *empty = full + MID_STAGES;
    uint64_t *acc_full2 = empty + MID_STAGES;
    uint64_t *acc_empty2 = acc_full2 + 3;
    uint64_t *w_full = acc_empty2 + 3;
    uint64_t *w1_full = w_full + 1;                                // layer 1's taps have landed (bulk copy, once per frame)
    float *s_par2 = reinterpret_cast<float *>(smem + SS::OFF_PAR2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ResizePlanDev &plan = src.plan;
    const int H = p.H, P1w = p.P1w, RPF = p.P1h + 1;
    const int Hc = min(H, 3 * p.P1h + 1);
    const int slot_idx = blockIdx.x;                              // this CTA's frame slot of the layer-1 buffer
    const int n_frames_cta = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // Layer-1 tiles of ONE frame: its P1h pooled rows only.  (The continuous stream of conv1_fused_tc also walks the zero row between
    // frames; here a frame stands alone, so the stream ends with the last real pooled row -- 32 tiles instead of 33 at 720p -- and
    // the unfold stops after the one zero row below the image that this row's views read.  Positions past it in the operand ring
    // hold stale data, which only reaches accumulator rows that are never stored.)
    const int n_tiles = (p.P1h * P1w + 127) / 128;
    const int total_u = 3 * p.P1h + 1;
    const int tiles2 = p2.FP / 128;                               // layer-2 tiles of one frame (the frame pitch is a multiple of 128)
    const int n_slots = src.n_slots, slot_bytes = src.n_src * src.row_bytes;    // (slot_bytes: one ROW's bytes in the ring)
    const uint32_t inv_slots = 0xffffffffu / (uint32_t)n_slots + 1u;
    // Source rows by TMA (src.tma_shift >= 0): a slot of the raw ring is fetched by ONE box of src_map -- 2^tma_shift rows of an
    // integer-scale gather or a plain copy (their rows are evenly spaced: the map's row axis walks them), or the two adjacent rows a
    // bilinear / 2x2 resize reads per output row (n_src = 2, tma_shift = 0: the map's row axis is the buffer's).  Why: a thread gets a bulk copy accepted every ~500 cycles and no faster
    // (300 cycles for the instruction alone on an idle SM, profiles/r02_timeline_frames_detail.txt), so three loaders issuing one
    // row each delivered 4.5 rows per ~1,400 cycles -- precisely the tile period of layer 1, whose unfold warps were seen to pick
    // every row up the moment it landed.  With two rows per instruction the row supply has a factor 2 in hand.
    const bool use_tma = src.tma_shift >= 0;
    const int rps_shift = use_tma ? src.tma_shift : 0;

    // ---- once per launch: what survives the phases
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        s_par[i] = p.par16[i];
        s_par2[i] = p2.bias[i]; s_par2[C + i] = p2.scale[i]; s_par2[2 * C + i] = p2.shift[i];
    }
    for (int i = threadIdx.x; i < UNFOLD_WARPS * F1_CMP_STRIDE; i += blockDim.x) s_cmp[i] = 0u;
    for (int y = threadIdx.x; y < H; y += blockDim.x) {
        int r0, r1, b0 = 2048, b1 = 0;
        if (plan.gather_step_x > 0) { r0 = r1 = plan.gather_off_y + y * plan.gather_step_y; }
        else if (plan.mode == RESIZE_COPY) { r0 = r1 = y; }
        else if (plan.mode == RESIZE_AREA2) { r0 = 2 * y; r1 = 2 * y + 1; }
        else { r0 = plan.y0[y]; r1 = plan.y1[y]; b0 = plan.b0[y]; b1 = plan.b1[y]; }
        if (src.compact) { r0 = plan.row_slot[r0]; r1 = (plan.mode == RESIZE_LINEAR && b1 == 0) ? r0 : plan.row_slot[r1]; }
        s_rowoff[2 * y] = (int)(r0 * src.row_pitch);
        // (with two-row TMA boxes the second row needs no offset of its own: the slot keeps the first row's INDEX in the buffer)
        s_rowoff[2 * y + 1] = (use_tma && src.n_src == 2) ? r0 : (int)(r1 * src.row_pitch);
        s_yb[2 * y] = b0;
        s_yb[2 * y + 1] = b1;
    }
    if (plan.mode == RESIZE_LINEAR && plan.gather_step_x == 0)
        for (int x = threadIdx.x; x < plan.dst_w; x += blockDim.x) {
            const int x0 = plan.x0[x], x1 = plan.x1[x];
            int a0 = plan.a0[x], a1 = plan.a1[x];
            if (x1 != x0 + 1) { a0 += a1; a1 = 0; }
            s_xtab[x] = make_int4(3 * x0, a0 | (a1 << 16), 0, 0);
        }
    if (threadIdx.x == 0) {
        for (int s = 0; s < MID_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 3); }      // three issuers free a stage
        for (int d = 0; d < 3; ++d) { mbar_init(&acc_full2[d], 1); mbar_init(&acc_empty2[d], EPI_WARPS); }
        mbar_init(w_full, 1);
        mbar_init(w1_full, 1);
        fence_barrier_init();
        tma_prefetch_desc(&in_map);
        if (use_tma) tma_prefetch_desc(&src_map);
    }
    if (warp == F1_MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    long long *tl = (p.timeline && blockIdx.x == 0 && threadIdx.x == 0) ? p.timeline : nullptr;    // debug stamps of CTA 0
    // ... and of its other roles during its THIRD frame (lane 0 of a warp): loaders [40 + row], unfold warp 0 [256 + 4 row ..],
    // issuer [700 + 2 tile: rows there, + 1: issued], epilogue [800 + tile: stored]
    long long *tlw = (p.timeline && blockIdx.x == 0 && lane == 0) ? p.timeline : nullptr;
    if (tl) tl[0] = clock64();
    if (p.timeline && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); p.timeline[2048 + 2 * blockIdx.x] = g; }
    grid_dep_launch();

    // ---- phase changes (every thread runs all three, once per frame)
    // Layer 1 starts from zero: its taps and the zero row above the frame go back over what layer 2 left, its counters and barriers
    // are reset.  (All of them are quiescent: every transaction was waited for by a consumer, every generic arrival precedes the
    // __syncthreads of frame_end, and the issuer waits for its last commit.)
    const uint32_t Z2 = 0u, Z2K = 0x3c000000u;
    // layer 1's barriers: acc_full[3] | acc_empty[3] | raw_full | raw_empty are consecutive, tile_done follows the counters
    constexpr int L1_BARS = 6 + 2 * RAW_SLOTS_MAX + TILE_RING;
    auto l1_bar = [&](int i) { return i < L1_BARS - TILE_RING ? acc_full + i : tile_done + (i - (L1_BARS - TILE_RING)); };
    auto frame_begin = [&](int it) {
        if (threadIdx.x == 0) {            // layer 1's taps by bulk copy (the issuer waits for them before its first MMA)
            mbar_arrive_expect_tx(w1_full, S::W_BYTES);
            bulk_load_1d(s_w, p.w_perm16, S::W_BYTES, w1_full);
        }
        for (int i = threadIdx.x; i < 2 * P1w; i += blockDim.x)
            reinterpret_cast<uint4 *>(s_ring + 2 * RG::SUB + (i >= P1w ? RG::PLANE : 0))[CAP - P1w + (i >= P1w ? i - P1w : i)] =
                make_uint4(Z2, Z2, Z2, i >= P1w ? Z2K : Z2);
        if (threadIdx.x < UNFOLD_WARPS + LOADER_WARPS) s_rows_done[threadIdx.x] = 0;    // ... and s_rows_issued
        {   // one barrier per thread (they were invalidated at the last phase change, see phase_switch)
            const int i = (int)threadIdx.x - 64;
            if (i >= 0 && i < L1_BARS) {
                const bool is_raw_empty = i >= 6 + RAW_SLOTS_MAX && i < 6 + 2 * RAW_SLOTS_MAX;
                mbar_init(l1_bar(i), (i >= 3 && i < 6) ? EPI_WARPS : (is_raw_empty ? 1 << rps_shift : 1));
                fence_barrier_init();
            }
        }
        if (tl && it < 600) tl[1 + 3 * it] = clock64();
        fence_proxy_async();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    };
    // layer 1 -> layer 2: the slot is complete in global memory (the epilogue warps fenced their stores towards the async proxy)
    auto phase_switch = [&](int it) {
        fence_proxy_async();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
        // Layer 1's barriers are invalidated HERE, a whole layer-2 phase before frame_begin initialises them again: mbarrier.inval
        // (SYNCS.CCTL.IV) directly followed by mbarrier.init (SYNCS.EXCH) of the same barrier by the same thread left barriers
        // invalid -- the kernel hung on its second frame -- while a version in which one thread invalidated all 62 and then
        // initialised all 62 ran: the invalidation is not ordered with an exchange that follows it closely.
        const int i = (int)threadIdx.x - 64;
        if (i >= 0 && i < L1_BARS) mbar_inval(l1_bar(i));
        if (tl && it < 600) tl[2 + 3 * it] = clock64();
    };
    auto frame_end = [&](int it) {
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
        if (tl && it < 600) tl[3 + 3 * it] = clock64();
    };

    // Layer 1's MMAs over a frame's tiles (conv1_fused_tc_kernel's issuer loop).
    auto layer1_issue = [&](int it) {
        const uint32_t w_addr = smem_u32(s_w), ring_addr = smem_u32(s_ring);
        const uint32_t idesc1 = instr_desc_f16_acc16(128, 3 * C);
        uint32_t acc_phase = 0;
        int r_hi = 127 / P1w, r_rem = 127 % P1w;
        mbar_wait(w1_full, it & 1);
        for (int t = 0; t < n_tiles; ++t) {
            // rows of sub-rings 1 and 2 up to pooled row (128 t + 127) / P1w, of sub-ring 0 one pooled row further
            const int u_hi = min(total_u - 1, 3 * (r_hi + 1));
            r_rem += 128;
#pragma unroll
            for (int k = 0; k < 3; ++k) { const bool c = r_rem >= P1w; r_rem -= c ? P1w : 0; r_hi += c ? 1 : 0; }
            const int mine = lane % UNFOLD_WARPS;
            const int need = u_hi >= mine ? (u_hi - mine) / UNFOLD_WARPS + 1 : 0;
            while (!__all_sync(0xffffffffu, ld_acquire_shared(&s_rows_done[mine]) >= need)) __nanosleep(32);
            __syncwarp();
            tc_fence_after_sync();
            if (tlw && it == 2 && t < 40) tlw[700 + 2 * t] = clock64();
            uint32_t a_chunk[5];
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const int sub = (c + 2) % 3, shift = c == 0 ? -P1w : (c == 4 ? P1w : 0);
                a_chunk[c] = ring_addr + sub * RG::SUB + (uint32_t)RG::wrap(t * 128 + shift + CAP) * 16;
            }
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                mbar_wait(&acc_empty[dy], acc_phase ^ 1);
                tc_fence_after_sync();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) {
                        const uint64_t da = smem_desc(a_chunk[dy + ks], RG::PLANE, 128);
                        const uint64_t db = smem_desc(w_addr + 2 * ks * S::LBO_B, S::LBO_B, 128);
                        umma_16bit(tmem_base + 3 * C * dy, da, db, idesc1, ks > 0 ? 1u : 0u);
                    }
                    umma_commit(&acc_full[dy]);
                    if (dy == 2) umma_commit(&tile_done[t & (TILE_RING - 1)]);   // tile t no longer reads the operand ring
                }
                __syncwarp();
            }
            if (tlw && it == 2 && t < 40) tlw[701 + 2 * t] = clock64();
            acc_phase ^= 1;
        }
        // the last commit has arrived (and with it every earlier one) before the barriers are invalidated
        mbar_wait(&tile_done[(n_tiles - 1) & (TILE_RING - 1)], ((n_tiles - 1) / TILE_RING) & 1);
    };

    // Layer 2's MMAs of ONE block row dy over a frame's tiles.  Three warps issue, one per block row (the layer-1 issuer and the two
    // loaders that layer 2 does not need): a block row is 15 MMAs per k-step behind ~300 instructions of descriptor set-up, which
    // one warp issues about as fast as the tensor pipe retires them; three in parallel keep its queue full.  The rows write
    // disjoint TMEM columns, each issuer commits its own row's barrier, and a stage is free once all three have committed it.
    const uint32_t w2_addr = smem_u32(s_w2), stage_addr = smem_u32(s_stage);
    auto layer2_issue = [&](auto dyc, int it, uint32_t &stage, uint32_t &phase, uint32_t &acc_phase2) {
        constexpr int dy = decltype(dyc)::value;
        mbar_wait(w_full, it & 1);
        for (int t2 = 0; t2 < tiles2; ++t2) {
            for (int ks = 0; ks < KS; ++ks) {
                mbar_wait(&full[stage], phase);
                if (ks == 0) mbar_wait(&acc_empty2[dy], acc_phase2 ^ 1);     // the epilogue is done with this row of the previous tile
                tc_fence_after_sync();
                const uint32_t a_stage = stage_addr + stage * MID_STAGE_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int oy = dy + ky - 1;
                        const int py = (oy + 3) % 3, sy = oy < 0 ? -1 : (oy > 2 ? 1 : 0);
#pragma unroll
                        for (int o = 0; o < 5; ++o) {
                            const int ox = (o == 0) ? 1 : (o == 1) ? 0 : (o == 2) ? 2 : (o == 3) ? -1 : 3;   // centre first
                            const int px = (ox + 3) % 3, sx = ox < 0 ? -1 : (ox > 2 ? 1 : 0);
                            const int dx_lo = ox - 1 < 0 ? 0 : ox - 1, dx_hi = ox + 1 > 2 ? 2 : ox + 1;
                            const int n_dx = dx_hi - dx_lo + 1, kx_start = ox + 1 - dx_lo;
                            const uint32_t idesc = instr_desc_16bit(128, C * n_dx, kBf16);
                            const uint32_t a_view = a_stage + (py * 3 + px) * (2 * MID_WIN * 16) + (uint32_t)(MID_HALO + sy * p2.PW + sx) * 16;
                            const uint32_t b_tap = w2_addr + ky * S2::W_KY_BYTES + 2 * ks * S2::LBO_B + (2 - kx_start) * C * 16;
                            const uint64_t da = smem_desc(a_view, MID_WIN * 16, 128);
                            const uint64_t db = smem_desc(b_tap, S2::LBO_B, 128);
                            umma_16bit(tmem_base + C * (dy * 3 + dx_lo), da, db, idesc, (ks == 0 && ky == 0 && ox == 1) ? 0u : 1u);
                        }
                    }
                    if (ks == KS - 1) umma_commit(&acc_full2[dy]);
                    umma_commit(&empty[stage]);        // one of the three arrivals that free the stage
                }
                __syncwarp();
                if (++stage == MID_STAGES) { stage = 0; phase ^= 1; }
            }
            acc_phase2 ^= 1;
        }
    };

    if (warp < 8) reg_alloc<RL::REGS_EPI>();
    else if (warp >= RL::WARPS - 4) reg_dealloc<RL::REGS_LIGHT>();
    else reg_dealloc<RL::REGS_UNFOLD>();
    if (warp >= F1_MMA_WARP) {
        if (warp >= F1_LOAD_WARP0) {
            // ------------------------------------------------------------------ loaders (layer 1); the first one is layer 2's TMA producer
            const int lw = warp - F1_LOAD_WARP0;
            const int n_loaders = min(LOADER_WARPS, n_slots);
            const uint64_t stream_once = l2_policy_evict_first();
            uint32_t stage = 0, phase = 0, acc_phase2 = 0;
            for (int it = 0; it < n_frames_cta; ++it) {
                frame_begin(it);
                if (lw < n_loaders) {
                    if (use_tma) {
                        // group g = rows [g << rps_shift, (g + 1) << rps_shift) of the frame: one TMA box, one slot, one barrier phase
                        const int n_groups = (Hc + (1 << rps_shift) - 1) >> rps_shift, f_idx = blockIdx.x + it * gridDim.x;
                        int issued = 0;
                        for (int g = lw; g < n_groups; g += n_loaders) {
                            const int use = (int)__umulhi((uint32_t)g, inv_slots), slot = g - use * n_slots;
                            const int row_coord = src.n_src == 2 ? s_rowoff[2 * g + 1] : g << rps_shift;
                            // (experiment) Where slots are scarce the loader waits here for thousands of cycles, and the box then takes a DRAM
                            // round trip (~2,300 cycles from issue to landing): ask for it NOW, so that the load below finds it in the L2.
                            if (src.l2_prefetch && elect_one()) tma_prefetch_4d_hint(&src_map, 0, 0, row_coord, f_idx, stream_once);
                            mbar_wait(&raw_empty[slot], (use & 1) ^ 1);
                            ++issued;
                            if (tlw && it == 2 && (g << rps_shift) < 200) tlw[40 + (g << rps_shift)] = clock64();
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)(slot_bytes << rps_shift));
                                tma_load_4d_hint(s_raw + (slot << rps_shift) * slot_bytes, &src_map, &raw_full[slot], 0, 0,
                                                 row_coord, f_idx, stream_once);
                                st_release_shared(&s_rows_issued[lw], issued);   // raw_full[slot] is now in group g's phase
                            }
                            __syncwarp();
                        }
                    } else {
                        // one bulk copy per source row (two for a bilinear resize), row n issued by loader n % n_loaders
                        const uint8_t *frame = src.frames + (long long)(blockIdx.x + (long long)it * gridDim.x) * src.frame_stride;
                        int issued = 0;
                        for (int n = lw; n < Hc; n += n_loaders) {
                            const int use = (int)__umulhi((uint32_t)n, inv_slots), slot = n - use * n_slots;
                            mbar_wait(&raw_empty[slot], (use & 1) ^ 1);
                            ++issued;
                            if (tlw && it == 2 && n < 200) tlw[40 + n] = clock64();
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)slot_bytes);
                                for (int j = 0; j < src.n_src; ++j)
                                    bulk_load_1d_hint(s_raw + slot * slot_bytes + j * src.row_bytes, frame + s_rowoff[2 * n + j],
                                                      (uint32_t)src.row_bytes, &raw_full[slot], stream_once);
                                st_release_shared(&s_rows_issued[lw], issued);
                            }
                            __syncwarp();
                        }
                    }
                }
                phase_switch(it);
                if (lw == 0) {
                    fence_proxy_async_global();
                    if (elect_one()) {
                        mbar_arrive_expect_tx(w_full, S2::W_BYTES);
                        for (int ky = 0; ky < 3; ++ky)
                            bulk_load_1d(s_w2 + ky * S2::W_KY_BYTES, reinterpret_cast<const uint8_t *>(p2.w_packed) + ky * S2::W_KY_BYTES, S2::W_KY_BYTES, w_full);
                    }
                    __syncwarp();
                    for (int t2 = 0; t2 < tiles2; ++t2) {
                        const int tile = slot_idx * tiles2 + t2;
                        for (int ks = 0; ks < KS; ++ks) {
                            mbar_wait(&empty[stage], phase ^ 1);
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&full[stage], MID_STAGE_BYTES);
                                uint8_t *dst = s_stage + stage * MID_STAGE_BYTES;
                                for (int pl = 0; pl < 9; ++pl)
                                    tma_load_4d(dst + pl * (2 * MID_WIN * 16), &in_map, &full[stage], 0, 4 * tile - 1, 2 * ks, pl);
                            }
                            __syncwarp();
                            if (++stage == MID_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                } else if (lw == 1) {
                    layer2_issue(std::integral_constant<int, 1>{}, it, stage, phase, acc_phase2);
                } else {
                    layer2_issue(std::integral_constant<int, 2>{}, it, stage, phase, acc_phase2);
                }
                frame_end(it);
            }
        } else {
            // ------------------------------------------------------------------ MMA issuer of layer 1 and of layer 2's block row 0
            uint32_t stage = 0, phase = 0, acc_phase2 = 0;
            for (int it = 0; it < n_frames_cta; ++it) {
                frame_begin(it);
                layer1_issue(it);
                phase_switch(it);
                layer2_issue(std::integral_constant<int, 0>{}, it, stage, phase, acc_phase2);
                frame_end(it);
            }
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ unfold (layer 1); idle during layer 2
        for (int it = 0; it < n_frames_cta; ++it) {
            frame_begin(it);
            {
                F1Ctx cx{&p, &src, s_ring, s_raw, s_cmp, s_yb, s_xtab, raw_full, raw_empty, tile_done, s_rows_done, s_rows_issued,
                         RPF, Hc, total_u, n_slots, slot_bytes, min(LOADER_WARPS, n_slots), inv_slots,
                         (it == 2 && p.timeline && blockIdx.x == 0) ? p.timeline : nullptr, rps_shift};
                f1_unfold_role<C, GATHER, true, UNFOLD_WARPS, CAP>(cx, warp - 8, lane);
            }
            phase_switch(it);
            if (warp == 8 && it + 1 < n_frames_cta) {
                // these warps idle while layer 2 runs: bring the first source rows of the next frame into the L2, so that layer 1
                // starts on L2 hits instead of a DRAM round trip per loader
                const uint8_t *next = src.frames + (long long)(blockIdx.x + (long long)(it + 1) * gridDim.x) * src.frame_stride;
                for (int n = lane; n < min(Hc, 2 * n_slots); n += 32)      // (n_src = 2: the second row follows the first)
                    bulk_prefetch_l2(next + s_rowoff[2 * n], (uint32_t)(src.n_src * src.row_bytes));
            }
            frame_end(it);
        }
    } else {
        // ------------------------------------------------------------------ epilogue of both layers
        const int q = warp & 3, half = warp >> 2, m = q * 32 + lane, ch0 = half * CH;
        const uint32_t tmem_thread = tmem_base + ((uint32_t)(q * 32) << 16) + half * 3 * CH;     // layer 1: columns [dy][half][dx][CH]
        const uint32_t tmem_thread2 = tmem_base + ((uint32_t)(q * 32) << 16) + ch0;              // layer 2: columns [dy][dx][C]
        uint32_t acc_phase2 = 0;
        // the slot is read back by this CTA ~25 us after it is written: ask the L2 to keep it (the frames stream with evict_first)
        const uint64_t keep_in_l2 = l2_policy_evict_last();
        // The kernel before this one (conv3 of the previous group, and through it the previous launch of this kernel) may still be
        // using the buffers written from here on.
        grid_dep_wait();
        zero_pads(p.out, CG, slot_idx, slot_idx + 1, threadIdx.x, EPI_WARPS * 32);
        if (slot_idx > 0) {
            // ... and the MID_HALO positions before this slot, which tile 0's shifted views read: the tail of the previous slot's
            // padding (never anything but zero), zeroed here as well so that this CTA does not depend on when its neighbour starts
            uint4 *act = reinterpret_cast<uint4 *>(p.out.ptr);
            for (int i = threadIdx.x; i < 9 * CG * MID_HALO; i += EPI_WARPS * 32)
                act[(size_t)(i / MID_HALO) * p.out.gtot + (size_t)slot_idx * p.out.FP - MID_HALO + i % MID_HALO] = make_uint4(0, 0, 0, 0);
        }
        for (int it = 0; it < n_frames_cta; ++it) {
            frame_begin(it);
            {
                uint32_t acc_phase = 0;
                int X = m % P1w, Y = m / P1w;                          // position 128 t + m = Y * P1w + X of this frame
                EpiRow16<CH> bufA, bufB;
                mbar_wait(&acc_full[0], 0);
                tc_fence_after_sync();
                epi_issue_row16<C>(tmem_thread, 0, bufA);
                auto tile = [&](int t, EpiRow16<CH> &cur, EpiRow16<CH> &nxt) {
                    const bool valid = Y < p.P1h;
                    uint4 *dst = store_addr16<C>(p.out, slot_idx, Y, X, ch0);
                    X += 128;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { const bool c = X >= P1w; X -= c ? P1w : 0; Y += c ? 1 : 0; }
                    uint32_t v[CH / 2];
                    bool requested = epilogue_tile_pipelined16<C>(tmem_thread, acc_full, acc_empty, acc_phase, lane, t + 1 < n_tiles, cur, nxt, v);
                    epilogue_scale_shift16<C>(s_par, ch0, v);
                    if (!requested && mbar_test_wait(&acc_full[0], acc_phase ^ 1)) {
                        tc_fence_after_sync();
                        epi_issue_row16<C>(tmem_thread, 0, nxt);
                        requested = true;
                    }
                    if (valid) store_pixel16_hint<C>(dst, p.out.gtot, v, keep_in_l2);
                    if (!requested) {
                        mbar_wait(&acc_full[0], acc_phase ^ 1);
                        tc_fence_after_sync();
                        epi_issue_row16<C>(tmem_thread, 0, nxt);
                    }
                    if (tl && it == 2 && t < 40) tl[800 + t] = clock64();
                    acc_phase ^= 1;
                };
                for (int t = 0; t < n_tiles; t += 2) {
                    tile(t, bufA, bufB);
                    if (t + 1 < n_tiles) tile(t + 1, bufB, bufA);
                }
            }
            fence_proxy_async_global();            // this thread's stores of the slot, before layer 2's TMA reads them
            phase_switch(it);
            {
                const int f_out = blockIdx.x + it * gridDim.x;         // frame index inside this launch = slot of the layer-2 buffer
                zero_pads(p2.out, CG, f_out, f_out + 1, threadIdx.x, EPI_WARPS * 32);
                for (int t2 = 0; t2 < tiles2; ++t2) {
                    const int pos = t2 * 128 + m;
                    const int Y = pos / p2.PW, X = pos % p2.PW;
                    const bool valid = Y < p2.out_h && X < p2.out_w;
                    float v[CH];
                    epilogue_tile<C>(tmem_thread2, acc_full2, acc_empty2, acc_phase2, lane, s_par2, ch0, v);
                    if (valid) store_pixel<C>(p2.out, f_out, Y, X, ch0, v);
                    acc_phase2 ^= 1;
                }
            }
            frame_end(it);
        }
    }
    if (p.timeline && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); p.timeline[2049 + 2 * blockIdx.x] = g; }
    if (warp == F1_MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ input packing
// K1 for the tensor-core path: decoded frames -> x-unfolded 16-bit operands (pixel value * kPixelScale / 255, RGB order).
// One thread per (frame, output row, pooled column): five resized pixels, 32 bytes out.
__device__ __forceinline__ void resized_pixel(const ResizePlanDev &plan, const uint8_t *frame, int64_t row_pitch, int compact,
                                              int y, int x, int (&v)[3]) {
    if (plan.gather_step_x > 0) {       // pure gather (integer scale, e.g. 720p -> 256x144: src[5y+2][5x+2])
        const int sy = plan.gather_off_y + y * plan.gather_step_y;
        const int r = compact ? plan.row_slot[sy] : sy;
        const uint8_t *q = frame + (int64_t)r * row_pitch + 3 * (plan.gather_off_x + x * plan.gather_step_x);
        v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
    } else if (plan.mode == RESIZE_COPY) {
        const int r = compact ? plan.row_slot[y] : y;
        const uint8_t *q = frame + (int64_t)r * row_pitch + 3 * x;
        v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
    } else if (plan.mode == RESIZE_AREA2) {
        const int r0 = compact ? plan.row_slot[2 * y] : 2 * y, r1 = compact ? plan.row_slot[2 * y + 1] : 2 * y + 1;
        const uint8_t *q0 = frame + (int64_t)r0 * row_pitch + 6 * x, *q1 = frame + (int64_t)r1 * row_pitch + 6 * x;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (q0[c] + q0[3 + c] + q1[c] + q1[3 + c] + 2) >> 2;
    } else {
        const int x0 = plan.x0[x], x1 = plan.x1[x], a0 = plan.a0[x], a1 = plan.a1[x];
        const int y0 = plan.y0[y], y1 = plan.y1[y], b0 = plan.b0[y], b1 = plan.b1[y];
        const uint8_t *q0 = frame + (int64_t)(compact ? plan.row_slot[y0] : y0) * row_pitch;
        int s0[3], s1[3] = {0, 0, 0};
#pragma unroll
        for (int c = 0; c < 3; ++c) s0[c] = a0 * q0[3 * x0 + c] + (a1 ? a1 * q0[3 * x1 + c] : 0);
        if (b1) {
            const uint8_t *q1 = frame + (int64_t)(compact ? plan.row_slot[y1] : y1) * row_pitch;
#pragma unroll
            for (int c = 0; c < 3; ++c) s1[c] = a0 * q1[3 * x0 + c] + (a1 ? a1 * q1[3 * x1 + c] : 0);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = min(max((((b0 * (s0[c] >> 4)) >> 16) + ((b1 * (s1[c] >> 4)) >> 16) + 2) >> 2, 0), 255);
    }
}

__global__ void __launch_bounds__(256) preprocess_xin_kernel(ResizePlanDev plan, const uint8_t *__restrict__ frames,
                                                             int64_t frame_stride, int64_t row_pitch, int compact, int batch,
                                                             int P1w, uint4 *__restrict__ xin) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)batch * plan.dst_h * P1w;
    if (idx >= total) return;
    const int px = (int)(idx % P1w), y = (int)((idx / P1w) % plan.dst_h), b = (int)(idx / ((int64_t)P1w * plan.dst_h));
    const uint8_t *frame = frames + (int64_t)b * frame_stride;
    float e[16];
#pragma unroll
    for (int col = 0; col < 5; ++col) {
        const int x = 3 * px - 1 + col;
        int v[3] = {0, 0, 0};
        if (x >= 0 && x < plan.dst_w) resized_pixel(plan, frame, row_pitch, compact, y, x, v);
        constexpr float kU8Scale = kPixelScale / 255.f;     // 1 (bf16: 0..255) or 1/256 (fp16: exact, in [0,1))
        e[col * 3 + 0] = (float)v[2] * kU8Scale;      // BGR -> RGB
        e[col * 3 + 1] = (float)v[1] * kU8Scale;
        e[col * 3 + 2] = (float)v[0] * kU8Scale;
    }
    e[15] = 0.f;
    uint4 lo, hi;
    lo.x = pack2(e[0], e[1]); lo.y = pack2(e[2], e[3]); lo.z = pack2(e[4], e[5]); lo.w = pack2(e[6], e[7]);
    hi.x = pack2(e[8], e[9]); hi.y = pack2(e[10], e[11]); hi.z = pack2(e[12], e[13]); hi.w = pack2(e[14], e[15]);
    xin[idx * 2] = lo;
    xin[idx * 2 + 1] = hi;
}

// The float entry point: x float32 [B,3,H,W] in [0,1] -> x-unfolded 16-bit operands (x * kPixelScale).
__global__ void __launch_bounds__(256) pack_xin_f32_kernel(const float *__restrict__ x, int batch, int H, int W, int P1w,
                                                           uint4 *__restrict__ xin) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)batch * H * P1w;
    if (idx >= total) return;
    const int px = (int)(idx % P1w), y = (int)((idx / P1w) % H), b = (int)(idx / ((int64_t)P1w * H));
    const float *img = x + (int64_t)b * 3 * H * W + (int64_t)y * W;
    float e[16];
#pragma unroll
    for (int col = 0; col < 5; ++col) {
        const int xx = 3 * px - 1 + col;
        const bool ok = xx >= 0 && xx < W;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) e[col * 3 + ch] = ok ? img[(int64_t)ch * H * W + xx] * kPixelScale : 0.f;
    }
    e[15] = 0.f;
    uint4 lo, hi;
    lo.x = pack2(e[0], e[1]); lo.y = pack2(e[2], e[3]); lo.z = pack2(e[4], e[5]); lo.w = pack2(e[6], e[7]);
    hi.x = pack2(e[8], e[9]); hi.y = pack2(e[10], e[11]); hi.z = pack2(e[12], e[13]); hi.w = pack2(e[14], e[15]);
    xin[idx * 2] = lo;
    xin[idx * 2 + 1] = hi;
}

// Test hook: phase-split 16-bit activation -> float32 NCHW.
__global__ void unpack_phase_split_kernel(const uint16_t *__restrict__ act, int gtot, int FP, int PW, int batch, int C, int ph,
                                          int pw, float *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)batch * C * ph * pw;
    if (idx >= total) return;
    const int x = (int)(idx % pw), y = (int)((idx / pw) % ph), c = (int)((idx / ((int64_t)pw * ph)) % C);
    const int b = (int)(idx / ((int64_t)pw * ph * C));
    const int plane = (y % 3) * 3 + (x % 3);
    const size_t src = (((size_t)plane * (C / 8) + c / 8) * gtot + (size_t)b * FP + (y / 3) * PW + x / 3) * 8 + c % 8;
    out[idx] = kBf16 ? __bfloat162float(__ushort_as_bfloat16(act[src])) : __half2float(__ushort_as_half(act[src]));
}

__global__ void unpack_plain_kernel(const float *__restrict__ act, int batch, int C, int npix, float *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)batch * C * npix;
    if (idx >= total) return;
    const int pix = (int)(idx % npix), c = (int)((idx / npix) % C), b = (int)(idx / ((int64_t)npix * C));
    out[idx] = act[((size_t)b * npix + pix) * C + c];
}

// ------------------------------------------------------------------------------------------------ batch-statistics BatchNorm
// Training-mode forward on the tensor-core path (reference training_scripts/learn_contrasts.py:100-107: modules never put in
// .eval()): a conv layer runs with the identity affine, which leaves relu(max + bias) in the phase-split buffer; these kernels
// take the per-channel sums over the batch (entries that are not pixels are zero and add nothing), turn them into
// scale = gamma / sqrt(var_biased + eps), shift = beta - mean * scale, and apply that to the real entries in place.
__global__ void __launch_bounds__(256) bn_ps_stats_kernel(const uint4 *__restrict__ act, int CG, int gtot, int n_pos, double *__restrict__ sums) {
    const int pc = blockIdx.y, cg = pc % CG;                 // pc = plane * CG + channel group; n_pos = frames * FP (the tail up to gtot is unused)
    float s[8], ss[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) s[c] = ss[c] = 0.f;
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n_pos; pos += gridDim.x * blockDim.x) {
        const uint4 q = act[(size_t)pc * gtot + pos];
        const __half2 *h = reinterpret_cast<const __half2 *>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            s[2 * j] += f.x; ss[2 * j] = fmaf(f.x, f.x, ss[2 * j]);
            s[2 * j + 1] += f.y; ss[2 * j + 1] = fmaf(f.y, f.y, ss[2 * j + 1]);
        }
    }
    __shared__ float red[2][8][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float a = s[c], b = ss[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (lane == 0) { red[0][c][warp] = a; red[1][c][warp] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        const int k = threadIdx.x >> 3, c = threadIdx.x & 7;
        double t = 0.0;
        for (int w8 = 0; w8 < 8; ++w8) t += (double)red[k][c][w8];
        atomicAdd(&sums[(cg * 8 + c) * 2 + k], t);
    }
}

__global__ void bn_finalize_kernel(double *__restrict__ sums, double count, const float *__restrict__ gamma, const float *__restrict__ beta,
                                   float eps, int C, float *__restrict__ scale, float *__restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = sums[2 * c] / count, var = fmax(sums[2 * c + 1] / count - mean * mean, 0.0);
    const float sc = gamma[c] * (float)(1.0 / sqrt(var + (double)eps));
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    sums[2 * c] = 0.0; sums[2 * c + 1] = 0.0;                // ready for the next layer
}

__global__ void __launch_bounds__(256) bn_ps_apply_kernel(uint4 *__restrict__ act, int CG, int gtot, int n_frames, int FP, int PW, int out_h,
                                                          int out_w, const float *__restrict__ scale, const float *__restrict__ shift) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x, pc = blockIdx.y;
    if (pos >= gtot) return;
    const int plane = pc / CG, cg = pc % CG;
    const int f = pos / FP, r = pos % FP, y = 3 * (r / PW) + plane / 3, x = 3 * (r % PW) + plane % 3;
    if (f >= n_frames || y >= out_h || x >= out_w) return;   // not a pixel: stays zero (the next conv's padding)
    uint4 q = act[(size_t)pc * gtot + pos];
    __half2 *h = reinterpret_cast<__half2 *>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 v = __half22float2(h[j]);
        const int c = cg * 8 + 2 * j;
        const uint32_t packed = pack2(fmaf(v.x, scale[c], shift[c]), fmaf(v.y, scale[c + 1], shift[c + 1]));
        h[j] = *reinterpret_cast<const __half2 *>(&packed);
    }
    act[(size_t)pc * gtot + pos] = q;
}

// ------------------------------------------------------------------------------------------------ head, first FC
// AdaptiveAvgPool2d + flatten + Linear folded into one [n_feat x 32] matrix (the pool is linear), then ReLU and the
// BatchNorm1d affine: out[f][o] = act(sum_k act3[f][k] * W[k][o] + bias[o]).  A block takes 16 frames; its HEAD_KG groups
// (64 threads each) walk alternate k slabs of HEAD_KT staged through shared memory (activations read coalesced along k
// and stored transposed, so that two frames are one 64-bit load); thread (fy, ox) of a group accumulates a 2 frame x
// 4 output register tile; the groups are summed through shared memory at the end.  A block's time is the chain of its
// groups' slabs (global-load latency + the slab's FMAs), not bandwidth: hence four short chains instead of two, and the next
// slab's loads are in flight in registers while the current one is consumed.
constexpr int HEAD_FRAMES = 16, HEAD_KT = 48, HEAD_KG = 4, HEAD_THREADS = 64 * HEAD_KG, HEAD_PITCH = HEAD_FRAMES + 4;

__global__ void __launch_bounds__(HEAD_THREADS) head_fc1_kernel(const float *__restrict__ act3, const float *__restrict__ w_folded,
                                                                const float *__restrict__ bias, const float *__restrict__ scale,
                                                                const float *__restrict__ shift, int batch, int n_feat, int hidden,
                                                                int relu, float *__restrict__ out) {
    __shared__ __align__(16) float s_a[HEAD_KG][HEAD_KT * HEAD_PITCH];    // [group][k][frame]
    __shared__ __align__(16) float s_w[HEAD_KG][HEAD_KT * 32];            // [group][k][out]
    static_assert((HEAD_KG - 1) * 64 * 8 <= HEAD_KG * HEAD_KT * 32, "the partial sums reuse s_w");
    const int grp = threadIdx.x >> 6, tid = threadIdx.x & 63, fy = tid >> 3, ox = tid & 7;   // frames 2*fy.., outputs 4*ox..
    const int f0 = blockIdx.x * HEAD_FRAMES;
    float *sa = s_a[grp], *sw = s_w[grp];
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    constexpr int NA = HEAD_FRAMES * HEAD_KT / 64, NW = HEAD_KT * 8 / 64;
    static_assert(NA * 64 == HEAD_FRAMES * HEAD_KT && NW * 64 == HEAD_KT * 8, "a slab is a whole number of loads per thread");
    float ta[NA];
    float4 tw[NW];
    // all the global loads of a slab (they are independent) into registers
    auto fetch = [&](int k0) {
        const int kn = min(HEAD_KT, n_feat - k0);
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const int i = tid + 64 * j, f = i / HEAD_KT, k = i % HEAD_KT;          // coalesced along k
            ta[j] = (f0 + f < batch && k < kn) ? __ldg(&act3[(size_t)(f0 + f) * n_feat + k0 + k]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const int i = tid + 64 * j, k = i >> 3;
            tw[j] = k < kn ? __ldg(&reinterpret_cast<const float4 *>(w_folded)[(size_t)(k0 + k) * 8 + (i & 7)]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    int k0 = grp * HEAD_KT;
    if (k0 < n_feat) fetch(k0);
    for (; k0 < n_feat; k0 += HEAD_KG * HEAD_KT) {
        bar_sync_named(1 + grp, 64);                  // the previous slab has been consumed
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const int i = tid + 64 * j;
            sa[(i % HEAD_KT) * HEAD_PITCH + i / HEAD_KT] = ta[j];
        }
#pragma unroll
        for (int j = 0; j < NW; ++j) reinterpret_cast<float4 *>(sw)[tid + 64 * j] = tw[j];
        bar_sync_named(1 + grp, 64);
        if (k0 + HEAD_KG * HEAD_KT < n_feat) fetch(k0 + HEAD_KG * HEAD_KT);      // in flight while this slab is consumed
#pragma unroll 8
        for (int k = 0; k < HEAD_KT; ++k) {
            const float2 a = *reinterpret_cast<const float2 *>(&sa[k * HEAD_PITCH + 2 * fy]);
            const float4 w = *reinterpret_cast<const float4 *>(&sw[k * 32 + 4 * ox]);
            acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
            acc[0][2] = fmaf(a.x, w.z, acc[0][2]); acc[0][3] = fmaf(a.x, w.w, acc[0][3]);
            acc[1][0] = fmaf(a.y, w.x, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
            acc[1][2] = fmaf(a.y, w.z, acc[1][2]); acc[1][3] = fmaf(a.y, w.w, acc[1][3]);
        }
    }
    __syncthreads();
    float *red = s_w[0];                              // [group - 1][64 threads][8]
    if (grp > 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[((grp - 1) * 64 + tid) * 8 + i * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (grp == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = f0 + 2 * fy + i;
            if (f >= batch) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = 4 * ox + j;
                if (o >= hidden) continue;
                float v = acc[i][j];
#pragma unroll
                for (int g = 0; g < HEAD_KG - 1; ++g) v += red[(g * 64 + tid) * 8 + i * 4 + j];
                v += bias[o];
                if (relu) v = fmaxf(v, 0.f);
                if (scale) v = fmaf(v, scale[o], shift[o]);
                out[(size_t)f * hidden + o] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ host state
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// A phase-split activation buffer as a 4-D tensor: (32 positions x 8 channels = 256 halves, block of 32 positions,
// channel group, phase plane); the box is one plane's window of a tile: 6 blocks x 2 channel groups.
// The source rows of `n_frames` frames as a 4-D tensor for conv12_frames' loaders: [frame][row][row_bytes / d0][d0 bytes], where the
// rows are the evenly spaced ones an integer-scale gather reads; a box is `rows_per_box` whole rows, contiguous in shared memory.
int make_src_map(CUtensorMap *map, const uint8_t *base, int row_bytes, long long row_stride, long long frame_stride, int n_rows,
                 int n_frames, int rows_per_box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(CUTDET_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    int d0 = 256;
    while (d0 > 16 && row_bytes % d0) d0 >>= 1;
    if (row_bytes % d0 || row_bytes / d0 > 256) return CUTDET_EUNSUPPORTED;
    cuuint64_t dims[4] = {(cuuint64_t)d0, (cuuint64_t)(row_bytes / d0), (cuuint64_t)n_rows, (cuuint64_t)n_frames};
    cuuint64_t strides[3] = {(cuuint64_t)d0, (cuuint64_t)row_stride, (cuuint64_t)frame_stride};
    cuuint32_t box[4] = {(cuuint32_t)d0, (cuuint32_t)(row_bytes / d0), (cuuint32_t)rows_per_box, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CUTDET_EUNSUPPORTED;      // a shape the encoder refuses: the caller stays on bulk copies
    return CUTDET_OK;
}

int make_act_map(CUtensorMap *map, void *base, int CG, int gtot) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(CUTDET_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {256, (cuuint64_t)gtot / 32, (cuuint64_t)CG, 9};
    cuuint64_t strides[3] = {512, (cuuint64_t)gtot * 16, (cuuint64_t)gtot * 16 * CG};
    cuuint32_t box[4] = {256, MID_WIN / 32, 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, kBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CUTDET_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CUTDET_OK;
}

}  // namespace

struct TcState {
    int C = 0;
    void *d_w1 = nullptr, *d_w2 = nullptr, *d_w3 = nullptr;   // packed 16-bit operands
    void *d_w1_perm = nullptr;                                 // conv1 taps, rows ordered [channel half][dx][C/2] (fused kernel)
    // conv1 with |BN scale| folded into its taps (see tc_prepare): epilogue parameters |s|*bias, sign(s), sign(s)/256
    float *d_c1_bias = nullptr, *d_c1_sign = nullptr, *d_c1_sign256 = nullptr;
    bool c1_folded = false;
    // ... and for the fused kernel with fp16 accumulators: taps * 2^a[co] (see tc_prepare), half2 pairs of +-2^(16-a) and shift
    void *d_w1_perm16 = nullptr;
    uint32_t *d_c1_par16 = nullptr;
    float *d_zero32 = nullptr;                                 // bias of the head when there is no FC layer
    // batch-statistics forward (training-mode BatchNorm): conv1 taps WITHOUT the folded BatchNorm scale, identity affine, scratch
    void *d_w1_plain = nullptr;
    float *d_ones = nullptr, *d_zeros = nullptr;               // [C]
    double *d_bn_sums = nullptr;                               // [C][2]
    float *d_bn_scale = nullptr, *d_bn_shift = nullptr;        // [C]
    std::map<std::tuple<const void *, int, int>, std::pair<CUtensorMap, CUtensorMap>> maps;   // (workspace, H*65536+W, sub) -> act1, act2 maps
    std::map<std::pair<int, int>, float *> fc1_folded;                                        // (P3h, P3w) -> [n_feat][32]
    std::mutex mutex;
};

namespace {

struct TcWorkspace {
    size_t xin, act1, act2, act3, fc[2], total;
    int sub, group_frames;      // frames per conv1/conv2 pass; frame capacity of the layer-2 buffer
    int gtot1, gtot2;
};

TcWorkspace tc_workspace(const cutdet_net *net, const Geom &g, int batch) {
    TcWorkspace w;
    const int sub_batch = net->opt.sub_batch > 0 ? net->opt.sub_batch : SUB_BATCH;
    const int group_frames = net->opt.group_frames > 0 ? net->opt.group_frames : GROUP_FRAMES;
    const int group_cap = std::max(group_frames, sub_batch);
    w.sub = batch < sub_batch ? batch : sub_batch;
    w.group_frames = batch < group_cap ? batch : group_cap;
    w.gtot1 = gtot_for(w.sub, g.FP1);
    w.gtot2 = gtot_for(w.group_frames, g.FP2);
    size_t off = 0;
    w.xin = off;  off = align_up(off + g.xin_frame * w.sub, 1024);
    w.act1 = off; off = align_up(off + act_bytes(g.CG, w.gtot1), 1024);
    w.act2 = off; off = align_up(off + act_bytes(g.CG, w.gtot2), 1024);
    w.act3 = off; off = align_up(off + g.act3_frame * batch, 1024);
    size_t widest = net->cfg.fc_hidden_size > net->cfg.fc_input_size ? net->cfg.fc_hidden_size : net->cfg.fc_input_size;
    for (int i = 0; i < 2; ++i) { w.fc[i] = off; off = align_up(off + widest * sizeof(float) * batch, 1024); }
    w.total = off;
    return w;
}

uint16_t operand_bits(float f) {      // float -> the 16-bit operand format, round to nearest even
    if (!kBf16) {
        const __half h = __float2half_rn(f);
        uint16_t b;
        memcpy(&b, &h, 2);
        return b;
    }
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

int upload_bytes(cutdet_net *net, const void *host, size_t bytes, void **dev) {
    CUTDET_CUDA(cudaMalloc(dev, bytes));
    net->dev_allocs.push_back(*dev);
    // A cudaMemcpy from pageable memory may return once the data is STAGED: the DMA then completes on the legacy stream, which
    // the caller's (non-blocking) stream does not wait for.  Uploads are rare (set-up, or once per input size): drain the device, so
    // that a kernel launched next on any stream sees the data (found by tests/test_gpu_stress.py: a net whose first forward -- the
    // lazily folded first FC matrix -- ran on a side stream while another net kept the GPU busy returned wrong logits).
    CUTDET_CUDA(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
    CUTDET_CUDA(cudaDeviceSynchronize());
    return CUTDET_OK;
}

template <int C>
int set_smem_limits() {
    CUTDET_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1Smem<C>::total));
    constexpr int smem32 = F1Smem<C, F1Roles<false>::UNFOLD_WARPS>::total, smem16 = F1Smem<C, F1Roles<true>::UNFOLD_WARPS>::total;
    static_assert(smem32 <= 232448 && smem16 <= 232448, "227 KB of shared memory per CTA");
    CUTDET_CUDA(cudaFuncSetAttribute(conv1_fused_tc_kernel<C, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem32));
    CUTDET_CUDA(cudaFuncSetAttribute(conv1_fused_tc_kernel<C, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem32));
    CUTDET_CUDA(cudaFuncSetAttribute(conv1_fused_tc_kernel<C, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem16));
    CUTDET_CUDA(cudaFuncSetAttribute(conv1_fused_tc_kernel<C, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem16));
    CUTDET_CUDA(cudaFuncSetAttribute(conv1_fused_sets_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem16));
    CUTDET_CUDA(cudaFuncSetAttribute(conv1_fused_sets_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem16));
    CUTDET_CUDA(cudaFuncSetAttribute(conv_mid_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, MidSmem<C>::total));
    CUTDET_CUDA(cudaFuncSetAttribute(conv12_frames_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F12Smem<C>::total));
    CUTDET_CUDA(cudaFuncSetAttribute(conv12_frames_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F12Smem<C>::total));
    CUTDET_CUDA(cudaFuncSetAttribute(conv12_frames_kernel<C, false, FR_CAP_TWO_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, F12Smem<C>::total));
    return CUTDET_OK;
}

// Launch with programmatic stream serialization: the kernel may start while the one before it in the stream drains (see
// grid_dep_launch / grid_dep_wait in tc_common.cuh; every kernel launched this way orders its dependent accesses itself).
// Grids are at most one CTA per SM, so a waiting successor can never keep a predecessor's CTA from being scheduled.
// cutdet_net_set_option(CUTDET_OPT_NO_PDL) falls back to ordinary launches (pdl = false everywhere).
// An optional window of global memory whose lines the kernel's accesses mark as PERSISTING in the L2 (cudaLaunchAttributeAccessPolicyWindow)
struct L2Window { void *base = nullptr; size_t bytes = 0; };

template <typename... KArgs, typename... Args>
void launch_pdl_window(bool pdl, const L2Window &window, void (*kernel)(KArgs...), int grid, int threads, size_t smem, cudaStream_t stream,
                       Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (window.base && window.bytes) {
        attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[n].val.accessPolicyWindow.base_ptr = window.base;
        attr[n].val.accessPolicyWindow.num_bytes = window.bytes;
        attr[n].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
void launch_pdl(bool pdl, void (*kernel)(KArgs...), int grid, int threads, size_t smem, cudaStream_t stream, Args &&...args) {
    launch_pdl_window(pdl, L2Window{}, kernel, grid, threads, smem, stream, std::forward<Args>(args)...);
}

template <int C>
int launch_conv1(const Conv1Params &p, cudaStream_t stream) {
    const int grid = p.B < sm_count() ? p.B : sm_count();
    {
        KernelScope scope("conv1_tc", stream);
        conv1_tc_kernel<C><<<grid, 416, C1Smem<C>::total, stream>>>(p);
    }
    CUTDET_LAUNCH_CHECK("conv1_tc_kernel");
    return CUTDET_OK;
}

template <int C>
int launch_conv1_fused(const Conv1Params &p, const FusedSrc &src_in, cudaStream_t stream, bool pdl, bool acc32, int grid_cap, int variant) {
    const bool use_sets = variant == 1;        // experiment: two epilogue sets + an MMA issuer per block row (conv1_fused_sets_kernel)
    // grid_cap (CUTDET_OPT_CONV1_GRID) is a test hook: several frames per CTA, as on a part with fewer SMs than a sub-batch has frames
    const int grid = std::min(std::min(p.B, sm_count()), grid_cap > 0 ? grid_cap : 1 << 30);
    static const bool regs_ok = [] {
        const void *fns[6] = {(const void *)conv1_fused_tc_kernel<C, true, false>, (const void *)conv1_fused_tc_kernel<C, false, false>,
                              (const void *)conv1_fused_tc_kernel<C, true, true>, (const void *)conv1_fused_tc_kernel<C, false, true>,
                              (const void *)conv1_fused_sets_kernel<C, true>, (const void *)conv1_fused_sets_kernel<C, false>};
        bool ok = true;
        for (int i = 0; i < 6; ++i) {
            cudaFuncAttributes a{};
            cudaFuncGetAttributes(&a, fns[i]);
            const int want = i < 2 ? F1Roles<false>::REGS_START : (i < 4 ? F1Roles<true>::REGS_START : S2Roles::REGS_START);
            if (a.numRegs != want) {
                fprintf(stderr, "cutdet: conv1_fused_tc compiled with %d registers, the setmaxnreg budget assumes %d\n", a.numRegs, want);
                ok = false;
            }
        }
        return ok;
    }();
    if (!regs_ok) return CUTDET_EUNSUPPORTED;
    // fp16 accumulators (half the TMEM read traffic, the max/affine on channel pairs) unless the net was told to keep fp32 ones
    // (cutdet_net_set_option(CUTDET_OPT_CONV1_ACC32))
    const bool acc16 = !acc32 && p.w_perm16 != nullptr;
    FusedSrc src = src_in;
    const long long ring_bytes = raw_bytes(acc16 ? F1Roles<true>::UNFOLD_WARPS : F1Roles<false>::UNFOLD_WARPS);
    src.n_slots = (int)std::min<long long>(ring_bytes / ((long long)src.n_src * src.row_bytes), RAW_SLOTS_MAX);
    // Every slot must belong to ONE loader (row n goes to loader n % n_loaders and to slot n % n_slots): a loader's wait on
    // raw_empty sees one parity bit, and only its own earlier fill of that slot keeps it from running two phases ahead.
    src.n_slots -= src.n_slots % std::min(LOADER_WARPS, src.n_slots);
    constexpr int smem32 = F1Smem<C, F1Roles<false>::UNFOLD_WARPS>::total, smem16 = F1Smem<C, F1Roles<true>::UNFOLD_WARPS>::total;
    constexpr int thr32 = F1Roles<false>::THREADS, thr16 = F1Roles<true>::THREADS;
    {
        KernelScope scope("conv1_fused_tc", stream);
        // fast path: integer-scale gather whose last pooled column has its right neighbour inside the image (dst_w % 3 != 0)
        const bool gather = src.plan.gather_step_x > 0 && src.plan.dst_w % 3 != 0;
        // pdl: only when the kernel before this one is one of ours (conv2/conv3 of the previous sub-batch): the loaders read the
        // frames without waiting for it, so it must not be what produced them
        if (acc16 && use_sets && gather) launch_pdl(pdl, conv1_fused_sets_kernel<C, true>, grid, S2Roles::THREADS, smem16, stream, p, src);
        else if (acc16 && use_sets) launch_pdl(pdl, conv1_fused_sets_kernel<C, false>, grid, S2Roles::THREADS, smem16, stream, p, src);
        else if (gather && acc16) launch_pdl(pdl, conv1_fused_tc_kernel<C, true, true>, grid, thr16, smem16, stream, p, src);
        else if (gather) launch_pdl(pdl, conv1_fused_tc_kernel<C, true, false>, grid, thr32, smem32, stream, p, src);
        else if (acc16) launch_pdl(pdl, conv1_fused_tc_kernel<C, false, true>, grid, thr16, smem16, stream, p, src);
        else launch_pdl(pdl, conv1_fused_tc_kernel<C, false, false>, grid, thr32, smem32, stream, p, src);
    }
    CUTDET_LAUNCH_CHECK("conv1_fused_tc_kernel");
    return CUTDET_OK;
}

// Can K1 be fused into conv1 for these frames?  (16-byte aligned rows for the bulk copies, tables that fit, enough raw slots.)
bool fused_source(const cutdet_resize_plan *plan, const cutdet_frames *frames, const Geom &g, FusedSrc *out) {
    const ResizePlanDev &h = plan->host;
    const long long row_bytes = 3LL * h.src_w;
    if (h.dst_w > F_MAX_DST || h.dst_h > F_MAX_DST || g.P1w < 64) return false;
    if (row_bytes % 16 || frames->row_pitch % 16 || frames->frame_stride % 16 || reinterpret_cast<uintptr_t>(frames->frames_dev) % 16)
        return false;
    const int rows = frames->row_map_compact ? plan->n_rows : h.src_h;
    if ((long long)rows * frames->row_pitch >= (1LL << 31)) return false;
    const int n_src = (h.gather_step_x > 0 || h.mode == RESIZE_COPY) ? 1 : 2;
    const long long slot = n_src * row_bytes;
    const long long n_slots = std::min<long long>(raw_bytes(8) / slot, RAW_SLOTS_MAX);   // the smaller ring; launch_conv1_fused sets the final count
    if (n_slots < 2) return false;
    out->plan = h;
    out->frames = frames->frames_dev;
    out->frame_stride = frames->frame_stride;
    out->row_pitch = frames->row_pitch;
    out->compact = frames->row_map_compact;
    out->n_src = n_src;
    out->row_bytes = (int)row_bytes;
    out->n_slots = (int)n_slots;
    out->tma_shift = -1;
    out->l2_prefetch = 0;
    out->pair_rows = plan->pair_rows ? 1 : 0;
    out->buf_rows = rows;
    return true;
}

template <int C>
int launch_mid(const CUtensorMap &map, const MidParams &p, const char *name, cudaStream_t stream, bool pdl) {
    const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
    {
        KernelScope scope(name, stream);
        launch_pdl(pdl, conv_mid_tc_kernel<C>, grid, MID_THREADS, MidSmem<C>::total, stream, map, p);
    }
    CUTDET_LAUNCH_CHECK("conv_mid_tc_kernel");
    return CUTDET_OK;
}

MidParams mid_params(int n_frames, int FP, int PW, int out_h, int out_w) {
    MidParams p;
    memset(&p, 0, sizeof(p));
    p.n_frames = n_frames; p.FP = FP; p.PW = PW; p.out_h = out_h; p.out_w = out_w;
    p.n_tiles = (int)(((long long)n_frames * FP + 127) / 128);
    return p;
}

int get_maps(cutdet_net *net, const Geom &g, const TcWorkspace &w, char *ws, std::pair<CUtensorMap, CUtensorMap> **out) {
    TcState *tc = net->tc;
    std::lock_guard<std::mutex> lock(tc->mutex);
    auto key = std::make_tuple((const void *)ws, g.H * 65536 + g.W, w.sub * 65536 + w.group_frames);
    auto it = tc->maps.find(key);
    if (it == tc->maps.end()) {
        std::pair<CUtensorMap, CUtensorMap> m;
        if (int rc = make_act_map(&m.first, ws + w.act1, g.CG, w.gtot1)) return rc;
        if (int rc = make_act_map(&m.second, ws + w.act2, g.CG, w.gtot2)) return rc;
        it = tc->maps.emplace(key, m).first;
    }
    *out = &it->second;
    return CUTDET_OK;
}

Conv1Params conv1_params(cutdet_net *net, const Geom &g, const TcWorkspace &w, char *ws, int nb) {
    TcState *tc = net->tc;
    Conv1Params c1;
    memset(&c1, 0, sizeof(c1));
    c1.xin = reinterpret_cast<const uint4 *>(ws + w.xin);
    c1.B = nb; c1.H = g.H; c1.P1h = g.P1h; c1.P1w = g.P1w;
    c1.tiles_per_frame = (g.P1h * g.P1w + 127) / 128;
    c1.out = OutSpec{ws + w.act1, 0, w.gtot1, g.PW1, g.FP1, g.Q1h, 0, g.P1h, g.P1w};
    c1.w_packed = reinterpret_cast<const uint4 *>(tc->d_w1);
    c1.w_perm = reinterpret_cast<const uint4 *>(tc->d_w1_perm);
    c1.bias = tc->c1_folded ? tc->d_c1_bias : net->conv[0].d_bias;
    c1.scale = tc->c1_folded ? tc->d_c1_sign : net->conv[0].d_scale;
    c1.shift = net->conv[0].d_shift;
    c1.scale_magic = tc->d_c1_sign256;
    c1.w_perm16 = reinterpret_cast<const uint4 *>(tc->d_w1_perm16);
    c1.par16 = tc->d_c1_par16;
    c1.folded = tc->c1_folded ? 1 : 0;
    return c1;
}

MidParams conv2_params(cutdet_net *net, const Geom &g, const TcWorkspace &w, char *ws, int nb, int slot0) {
    MidParams p2 = mid_params(nb, g.FP1, g.PW1, g.P2h, g.P2w);
    p2.out = OutSpec{ws + w.act2, 0, w.gtot2, g.PW2, g.FP2, g.Q2h, slot0, g.P2h, g.P2w};
    p2.w_packed = reinterpret_cast<const uint4 *>(net->tc->d_w2);
    p2.bias = net->conv[1].d_bias; p2.scale = net->conv[1].d_scale; p2.shift = net->conv[1].d_shift;
    return p2;
}

// conv1 + conv2 over one sub-batch whose x-unfolded input already sits in the workspace; layer 2's maps land in the
// group buffer at frame slot `slot0`.
template <int C>
int run_conv12(cutdet_net *net, const Geom &g, const TcWorkspace &w, char *ws, const CUtensorMap &map1, int nb, int slot0,
               cudaStream_t stream, const FusedSrc *fused, int f0) {
    Conv1Params c1 = conv1_params(net, g, w, ws, nb);
    // debug stamps (cutdet_net_debug_timeline): the armed buffer is caller-owned device memory; nothing is allocated, copied or
    // synchronised here
    if (net->opt.timeline_kernel == 1 && nb >= SUB_BATCH) c1.timeline = net->opt.timeline_dev;
    if (fused) {
        FusedSrc fs = *fused;
        fs.frames += (long long)f0 * fs.frame_stride;
        if (int rc = launch_conv1_fused<C>(c1, fs, stream, f0 > 0 && !net->opt.no_pdl, net->opt.conv1_acc32 != 0, net->opt.conv1_grid, net->opt.conv1_variant)) return rc;
    } else if (int rc = launch_conv1<C>(c1, stream)) return rc;

    MidParams p2 = conv2_params(net, g, w, ws, nb, slot0);
    if (net->opt.timeline_kernel == 2 && nb >= SUB_BATCH) p2.timeline = net->opt.timeline_dev;
    return launch_mid<C>(map1, p2, "conv2_tc", stream, !net->opt.no_pdl);
}

// K1 + conv1 + conv2 of frames [f0, f0 + n) in ONE launch, a frame at a time per CTA (conv12_frames_kernel); layer 2's maps land in
// frame slots [0, n) of the group buffer.  The layer-1 buffer serves as one scratch slot per CTA.
template <int C>
int run_conv12_frames(cutdet_net *net, const Geom &g, const TcWorkspace &w, char *ws, const CUtensorMap &map1, int n, cudaStream_t stream,
                      const FusedSrc &fused, int f0) {
    Conv1Params c1 = conv1_params(net, g, w, ws, n);
    if (net->opt.timeline_kernel == 3 && n >= SUB_BATCH) c1.timeline = net->opt.timeline_dev;
    const MidParams p2 = conv2_params(net, g, w, ws, n, 0);
    FusedSrc src = fused;
    src.frames += (long long)f0 * src.frame_stride;
    const int grid_cap = net->opt.conv1_grid > 0 ? net->opt.conv1_grid : 1 << 30;
    const int grid = std::min(std::min(std::min(n, sm_count()), w.sub), grid_cap);      // one layer-1 slot per CTA
    static const bool regs_ok = [] {
        const void *fns[3] = {(const void *)conv12_frames_kernel<C, true>, (const void *)conv12_frames_kernel<C, false>,
                              (const void *)conv12_frames_kernel<C, false, FR_CAP_TWO_ROWS>};
        bool ok = true;
        for (int i = 0; i < 3; ++i) {
            cudaFuncAttributes a{};
            cudaFuncGetAttributes(&a, fns[i]);
            if (a.numRegs != F1Roles<true>::REGS_START) {
                fprintf(stderr, "cutdet: conv12_frames compiled with %d registers, the setmaxnreg budget assumes %d\n", a.numRegs, F1Roles<true>::REGS_START);
                ok = false;
            }
        }
        return ok;
    }();
    if (!regs_ok) return CUTDET_EUNSUPPORTED;
    const bool gather = src.plan.gather_step_x > 0 && src.plan.dst_w % 3 != 0;
    // Two source rows per resized row (bilinear, 2x2): the raw ring's slots are what the unfold warps wait for (six 11.5 KB slots for
    // eight warps at 1080p), so the operand ring gives up a quarter (net option ring_cap: 1 = keep the full operand ring)
    // -- only where slots are scarce: 1080p 0.825 -> 0.706 ms per 1,184 frames with nine slots instead of six, 640x360 (21 slots
    // anyway) 1 % slower with the smaller operand ring (tools/ab_ring.py, profiles/r02_ab_ring.txt)
    const long long full_slots = raw_bytes(F1Roles<true>::UNFOLD_WARPS) / ((long long)src.n_src * src.row_bytes);
    const bool small_ring = !gather && src.n_src == 2 && (net->opt.ring_cap == 2 || (net->opt.ring_cap == 0 && full_slots < 2 * F1Roles<true>::UNFOLD_WARPS));
    const long long raw_ring = raw_bytes(F1Roles<true>::UNFOLD_WARPS) + (small_ring ? FrRing<FR_CAP_TWO_ROWS>::EXTRA_RAW : 0);
    src.n_slots = (int)std::min<long long>(raw_ring / ((long long)src.n_src * src.row_bytes), RAW_SLOTS_MAX);
    src.n_slots -= src.n_slots % std::min(LOADER_WARPS, src.n_slots);
    // Integer-scale gathers read evenly spaced source rows (row off_y + y * step_y; consecutive slots in compact frames): the
    // loaders then fetch two rows per TMA instruction through a tensor map of this launch's frames (see the kernel).
    CUtensorMap src_map;
    memset(&src_map, 0, sizeof(src_map));
    src.tma_shift = -1;
    const bool even_rows = (gather || src.plan.mode == RESIZE_COPY) && src.n_src == 1;
    if (!even_rows && src.n_src == 2 && src.pair_rows && !src.plan.gather_step_x) {
        // the two source rows of an output row as one two-row box at the first one's index in the buffer
        if (make_src_map(&src_map, src.frames, src.row_bytes, src.row_pitch, src.frame_stride, src.buf_rows, n, 2) == CUTDET_OK) src.tma_shift = 0;
    }
    if (even_rows) {
        const int shift = 1, rows_per_slot = 1 << shift;
        int groups = (int)std::min<long long>(raw_bytes(F1Roles<true>::UNFOLD_WARPS) / ((long long)rows_per_slot * src.row_bytes), RAW_SLOTS_MAX);
        groups -= groups % LOADER_WARPS;
        const bool plain = src.plan.mode == RESIZE_COPY && !gather;
        const long long row_stride = (src.compact || plain) ? src.row_pitch : (long long)src.plan.gather_step_y * src.row_pitch;
        const uint8_t *base = src.frames + ((src.compact || plain) ? 0 : (long long)src.plan.gather_off_y * src.row_pitch);
        if (groups >= LOADER_WARPS && c1.H % rows_per_slot == 0 &&
            make_src_map(&src_map, base, src.row_bytes, row_stride, src.frame_stride, c1.H, n, rows_per_slot) == CUTDET_OK) {
            src.tma_shift = shift;
            src.n_slots = groups;
        }
    }
    // CUTDET_OPT_SRC_PREFETCH (experiment, off by default): 1 = where a resized row reads two source rows (few, large slots: the
    // loaders wait for slots), 2 = with every tensor map of the source rows.  Measured (profiles/r02_ab_ring_prefetch.txt): 1080p with
    // SIX slots 0.822 -> 0.788 ms per 1,184 frames, with the nine slots of the smaller operand ring 0.706 -> 0.705 (the unfold warps'
    // resize arithmetic is the limit by then); 640x360 and 720p 0.3-1 % slower (one more instruction per box for the loaders).
    src.l2_prefetch = src.tma_shift >= 0 && (net->opt.src_prefetch == 2 || (net->opt.src_prefetch == 1 && src.n_src == 2)) ? 1 : 0;
    L2Window window;
    if (net->opt.l2_persist) {
        // experiment (CUTDET_OPT_L2_PERSIST): the layer-1 slots as a persisting window of the L2, so that the frame stream cannot push
        // dirty slot lines out to DRAM between their write and their read-back.  The carve-out is a limit of the CUDA context.
        static const size_t carve = [] {
            int dev = 0, max_persist = 0, max_window = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            if (max_persist > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
            return (size_t)std::min(max_persist, max_window);
        }();
        window = L2Window{ws + w.act1, std::min(carve, act_bytes(g.CG, w.gtot1))};
    }
    {
        KernelScope scope("conv12_frames", stream);
        const bool pdl = f0 > 0 && !net->opt.no_pdl;        // the loaders read the frames at once: only behind a kernel of ours
        if (gather) launch_pdl_window(pdl, window, conv12_frames_kernel<C, true>, grid, F1Roles<true>::THREADS, (size_t)F12Smem<C>::total, stream, c1, src, map1, p2, src_map);
        else if (small_ring) launch_pdl_window(pdl, window, conv12_frames_kernel<C, false, FR_CAP_TWO_ROWS>, grid, F1Roles<true>::THREADS, (size_t)F12Smem<C>::total, stream, c1, src, map1, p2, src_map);
        else launch_pdl_window(pdl, window, conv12_frames_kernel<C, false>, grid, F1Roles<true>::THREADS, (size_t)F12Smem<C>::total, stream, c1, src, map1, p2, src_map);
    }
    CUTDET_LAUNCH_CHECK("conv12_frames_kernel");
    return CUTDET_OK;
}

// conv3 over the `n` frames gathered in the group buffer; their layer-3 maps go to frames [frame0, frame0 + n) of act3.
template <int C>
int run_conv3(cutdet_net *net, const Geom &g, const TcWorkspace &w, char *ws, const CUtensorMap &map2, int n, int frame0,
              cudaStream_t stream) {
    TcState *tc = net->tc;
    MidParams p3 = mid_params(n, g.FP2, g.PW2, g.P3h, g.P3w);
    p3.out = OutSpec{ws + w.act3, 1, 0, 0, 0, 0, frame0, g.P3h, g.P3w};
    p3.w_packed = reinterpret_cast<const uint4 *>(tc->d_w3);
    p3.bias = net->conv[2].d_bias; p3.scale = net->conv[2].d_scale; p3.shift = net->conv[2].d_shift;
    return launch_mid<C>(map2, p3, "conv3_tc", stream, !net->opt.no_pdl);
}

// AdaptiveAvgPool + first FC folded, for this pooled-map size; built once and cached.
int folded_fc1(cutdet_net *net, const Geom &g, float **out) {
    TcState *tc = net->tc;
    std::lock_guard<std::mutex> lock(tc->mutex);
    auto key = std::make_pair(g.P3h, g.P3w);
    auto it = tc->fc1_folded.find(key);
    if (it != tc->fc1_folded.end()) { *out = it->second; return CUTDET_OK; }
    const int P = net->cfg.avg_pool_size, C = g.C, npix = g.P3h * g.P3w, n_feat = npix * C;
    // no FC layer: the "weights" are the identity on the flattened pooled features (c, i, j), so the fold is the pooling matrix
    const bool identity = net->cfg.n_fc_layers == 0;
    struct { int in, out; const float *w; } L;
    if (identity) { L.in = L.out = C * P * P; L.w = nullptr; }
    else { L.in = net->fc[0].in; L.out = net->fc[0].out; L.w = net->fc[0].w.data(); }
    std::vector<float> folded((size_t)n_feat * 32, 0.f);
    std::vector<double> acc((size_t)n_feat * L.out, 0.0);
    for (int i = 0; i < P; ++i) {
        const int r0 = (i * g.P3h) / P, r1 = ((i + 1) * g.P3h + P - 1) / P;
        for (int j = 0; j < P; ++j) {
            const int c0 = (j * g.P3w) / P, c1 = ((j + 1) * g.P3w + P - 1) / P;
            const double inv = 1.0 / ((r1 - r0) * (c1 - c0));
            for (int r = r0; r < r1; ++r)
                for (int cc = c0; cc < c1; ++cc)
                    for (int ch = 0; ch < C; ++ch) {
                        const int feat = ch * P * P + i * P + j;
                        if (identity) { acc[((size_t)(r * g.P3w + cc) * C + ch) * L.out + feat] += inv; continue; }
                        for (int o = 0; o < L.out; ++o)
                            acc[((size_t)(r * g.P3w + cc) * C + ch) * L.out + o] += inv * L.w[(size_t)o * L.in + feat];
                    }
        }
    }
    for (int f = 0; f < n_feat; ++f)
        for (int o = 0; o < L.out; ++o) folded[(size_t)f * 32 + o] = (float)acc[(size_t)f * L.out + o];
    void *d = nullptr;
    if (int rc = upload_bytes(net, folded.data(), folded.size() * sizeof(float), &d)) return rc;
    tc->fc1_folded[key] = reinterpret_cast<float *>(d);
    *out = reinterpret_cast<float *>(d);
    return CUTDET_OK;
}

int run_head(cutdet_net *net, const Geom &g, const TcWorkspace &w, char *ws, int batch, float *logits, cudaStream_t stream) {
    const float *cur = reinterpret_cast<const float *>(ws + w.act3);
    const int n_feat = g.P3h * g.P3w * g.C;
    float *folded = nullptr;
    if (int rc = folded_fc1(net, g, &folded)) return rc;
    if (net->cfg.n_fc_layers == 0) {
        // a bare FrameConvNet: the pooled features [B, C*P*P] in (c, i, j) order ARE the output (avg-pool as a matrix, no bias)
        KernelScope scope("head_fc1", stream);
        head_fc1_kernel<<<(batch + HEAD_FRAMES - 1) / HEAD_FRAMES, HEAD_THREADS, 0, stream>>>(
            cur, folded, net->tc->d_zero32, nullptr, nullptr, batch, n_feat, g.C * net->cfg.avg_pool_size * net->cfg.avg_pool_size, 0, logits);
        CUTDET_LAUNCH_CHECK("head_fc1_kernel");
        return CUTDET_OK;
    }
    const FcLayer &L0 = net->fc[0];
    const bool last0 = net->cfg.n_fc_layers == 1;
    float *out0 = last0 ? logits : reinterpret_cast<float *>(ws + w.fc[0]);
    {
        KernelScope scope("head_fc1", stream);
        head_fc1_kernel<<<(batch + HEAD_FRAMES - 1) / HEAD_FRAMES, HEAD_THREADS, 0, stream>>>(
            cur, folded, L0.d_bias, L0.has_bn ? L0.d_scale : nullptr, L0.has_bn ? L0.d_shift : nullptr, batch, n_feat, L0.out,
            last0 ? 0 : 1, out0);
    }
    CUTDET_LAUNCH_CHECK("head_fc1_kernel");
    cur = out0;
    for (int j = 1; j < net->cfg.n_fc_layers; ++j) {
        const bool is_last = j + 1 == net->cfg.n_fc_layers;
        float *out = is_last ? logits : reinterpret_cast<float *>(ws + w.fc[j & 1]);
        if (int rc = launch_fc(cur, out, net->fc[j], batch, !is_last, stream)) return rc;
        cur = out;
    }
    return CUTDET_OK;
}

// The whole stack over a batch: `pack(f0, nb)` writes the x-unfolded input of frames [f0, f0 + nb) into the workspace.
template <int C, typename Pack>
int run_batch(cutdet_net *net, const Geom &g, const TcWorkspace &w, char *ws, int batch, float *logits, cudaStream_t stream,
              Pack pack, const FusedSrc *fused = nullptr) {
    std::pair<CUtensorMap, CUtensorMap> *maps = nullptr;
    if (int rc = get_maps(net, g, w, ws, &maps)) return rc;
    // CUTDET_OPT_CONV1_VARIANT: 0 = automatic, 2 = the same choice stated; 1 and 3 keep the two-kernel path (3 with the default conv1)
    if (fused && (net->opt.conv1_variant == 0 || net->opt.conv1_variant == 2) && !net->opt.conv1_acc32 && net->tc->d_w1_perm16) {
        // both layers of a frame by one CTA, a whole group per launch, then conv3 of the group
        for (int f0 = 0; f0 < batch; f0 += w.group_frames) {
            const int n = std::min(batch - f0, w.group_frames);
            if (int rc = run_conv12_frames<C>(net, g, w, ws, maps->first, n, stream, *fused, f0)) return rc;
            if (int rc = run_conv3<C>(net, g, w, ws, maps->second, n, f0, stream)) return rc;
        }
        return run_head(net, g, w, ws, batch, logits, stream);
    }
    int group_start = 0, in_group = 0;
    for (int f0 = 0; f0 < batch; f0 += w.sub) {
        const int nb = batch - f0 < w.sub ? batch - f0 : w.sub;
        if (!fused)
            if (int rc = pack(f0, nb)) return rc;
        if (int rc = run_conv12<C>(net, g, w, ws, maps->first, nb, in_group, stream, fused, f0)) return rc;
        in_group += nb;
        if (in_group + w.sub > w.group_frames || f0 + nb >= batch) {
            if (int rc = run_conv3<C>(net, g, w, ws, maps->second, in_group, group_start, stream)) return rc;
            group_start += in_group;
            in_group = 0;
        }
    }
    return run_head(net, g, w, ws, batch, logits, stream);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ interface
bool tc_supported(const cutdet_net *net, int height, int width) {
    if (!net->tc) return false;
    const Geom g = make_geom(height, width, net->cfg.hidden_channels);
    if (g.P3h < 1 || g.P3w < 1) return false;
    if (g.PW1 + 1 > MID_HALO || g.PW2 + 1 > MID_HALO) return false;      // a shifted view must stay inside the tile's window
    return true;
}

int tc_prepare(cutdet_net *net) {
    const cutdet_net_config &c = net->cfg;
    if (c.n_conv_layers != 3 || c.input_channels != 3 || (c.hidden_channels != 48 && c.hidden_channels != 32)) return CUTDET_OK;
    // the head kernel produces up to 32 features per frame: the first FC layer's outputs, or -- for a bare FrameConvNet
    // (the contrastive encoder's trunk, learn_contrasts.py:68-70) -- the C * P * P pooled features themselves
    if (c.n_fc_layers == 0 ? c.hidden_channels * c.avg_pool_size * c.avg_pool_size > 32
                           : (c.fc_hidden_size > 32 || (c.n_fc_layers == 1 && c.fc_output_size > 32)))
        return CUTDET_OK;
    if (!encode_fn()) return CUTDET_OK;          // no TMA descriptor encoder in this driver: stay on the generic kernels
    const int C = c.hidden_channels, CG = C / 8;
    TcState *tc = new TcState();
    tc->C = C;
    net->tc = tc;
    {
        const std::vector<float> zeros(32, 0.f);
        if (int rc = upload_bytes(net, zeros.data(), zeros.size() * sizeof(float), reinterpret_cast<void **>(&tc->d_zero32))) return rc;
    }
    // conv1 B operand: [ky][half][n = dx*C + co][8], k16 = col*3 + ch, taps / 255.
    // |BN scale| is folded into the taps (s * relu(u) = sign(s) * relu(|s| * u)), so the epilogue is relu(max + |s|*bias) * sign + shift;
    // and K element 15, which no pixel uses, carries the bias for the fused gather path: there the pixels enter as 1024 + v and
    // that element is the constant 1024, so with taps w'' and W = sum of the 27 rounded taps,
    //     acc = sum w'' (1024 + v) + 1024 * (wb0 + wb1 + wb2) = 256 * (z'' + 4W) + 1024 * wb,   wb = (|s|*bias - 4W) / 4 in three fp16 pieces
    // = 256 * (z'' + |s|*bias): bias and ReLU need no instruction of their own.  (Other paths leave element 15 at zero.)
    {
        std::vector<uint16_t> w((size_t)6 * 3 * C * 8, 0);
        const ConvLayer &L = net->conv[0];
        bool fold = !kBf16;
        for (int co = 0; co < C && fold; ++co) {
            const float a = fabsf(L.scale[co]);
            if (!std::isfinite(a)) fold = false;
            for (int i = 0; i < 27 && fold; ++i)
                if (!(fabsf(L.w[(size_t)co * 27 + i] * kW1Scale * a) < 16384.f)) fold = false;
        }
        tc->c1_folded = fold;
        std::vector<float> pb(C), ps(C), ps256(C);
        for (int co = 0; co < C; ++co) {
            const float a = fold ? fabsf(L.scale[co]) : 1.f;
            double W = 0;
            for (int ky = 0; ky < 3; ++ky)
                for (int col = 0; col < 5; ++col)
                    for (int ch = 0; ch < 3; ++ch)
                        for (int dx = 0; dx < 3; ++dx) {
                            const int kx = col - dx;
                            if (kx < 0 || kx > 2) continue;
                            const int k16 = col * 3 + ch;
                            const uint16_t bits = operand_bits(L.w[((size_t)co * 3 + ch) * 9 + ky * 3 + kx] * kW1Scale * a);
                            w[(((size_t)(2 * ky + k16 / 8)) * 3 * C + dx * C + co) * 8 + k16 % 8] = bits;
                            if (dx == 0 && !kBf16) W += (double)__half2float(__ushort_as_half(bits));
                        }
            pb[co] = a * L.bias[co];
            ps[co] = L.scale[co] < 0.f ? -1.f : 1.f;
            ps256[co] = ps[co] / 256.f;
            if (fold) {
                double rest = ((double)pb[co] - 4.0 * W) / 4.0;
                for (int ky = 0; ky < 3; ++ky) {
                    const __half piece = __float2half_rn((float)rest);
                    rest -= (double)__half2float(piece);
                    for (int dx = 0; dx < 3; ++dx) w[(((size_t)(2 * ky + 1)) * 3 * C + dx * C + co) * 8 + 7] = __half_as_ushort(piece);
                }
            }
        }
        if (int rc = upload_bytes(net, w.data(), w.size() * 2, &tc->d_w1)) return rc;
        {   // the same rows in the order the fused kernel's epilogue reads its accumulator columns: [channel half][dx][C/2]
            const int CHh = C / 2;
            std::vector<uint16_t> wq(w.size(), 0);
            for (int kyh = 0; kyh < 6; ++kyh)
                for (int dx = 0; dx < 3; ++dx)
                    for (int co = 0; co < C; ++co)
                        for (int e = 0; e < 8; ++e)
                            wq[(((size_t)kyh) * 3 * C + (co / CHh) * 3 * CHh + dx * CHh + co % CHh) * 8 + e] = w[(((size_t)kyh) * 3 * C + dx * C + co) * 8 + e];
            if (int rc = upload_bytes(net, wq.data(), wq.size() * 2, &tc->d_w1_perm)) return rc;
        }
        if (fold) {
            // The fused kernel with fp16 accumulators.  Pixels enter as the fp16 subnormals v * 2^-24 (the byte itself under a zero
            // byte), taps as w'' * 2^a with a per channel such that the largest tap lands in [2^14, 2^15): acc = 2^(a-16) * (z'' +
            // |s| bias) is at most 27 * 2^15 * 255 * 2^-24 = 13.4, so no partial sum can overflow, and values down to 2^-24 * 2^(16-a)
            // are kept (fp16 subnormals).  The constant K element is 1.0 and multiplies the bias, split over the three kernel rows
            // in fp16 pieces (each refines the rest).  The epilogue multiplies by +-2^(16-a) (exact) and adds the shift.
            const int CHh = C / 2;
            std::vector<uint16_t> wq((size_t)6 * 3 * C * 8, 0);
            std::vector<uint32_t> par(C, 0);
            auto half_bits = [](float f) { const __half h = __float2half_rn(f); return __half_as_ushort(h); };
            bool ok = true;
            for (int co = 0; co < C && ok; ++co) {
                const float a = fabsf(L.scale[co]);
                float wmax = 0.f;
                for (int i = 0; i < 27; ++i) wmax = std::max(wmax, fabsf(L.w[(size_t)co * 27 + i] * kW1Scale * a));
                int e = 0;
                if (wmax > 0.f) { frexpf(wmax, &e); }                    // wmax = m * 2^e, m in [0.5, 1)
                const int ae = 15 - e;                                    // wmax * 2^ae in [2^14, 2^15)
                if (ae < 2 || ae > 30) { ok = false; break; }             // scale 2^(16-ae) and the pieces must be fp16 numbers
                const float up = ldexpf(1.f, ae);
                for (int ky = 0; ky < 3; ++ky)
                    for (int col = 0; col < 5; ++col)
                        for (int ch = 0; ch < 3; ++ch)
                            for (int dx = 0; dx < 3; ++dx) {
                                const int kx = col - dx, k16 = col * 3 + ch;
                                if (kx < 0 || kx > 2) continue;
                                const float v = L.w[((size_t)co * 3 + ch) * 9 + ky * 3 + kx] * kW1Scale * a;
                                wq[(((size_t)(2 * ky + k16 / 8)) * 3 * C + (co / CHh) * 3 * CHh + dx * CHh + co % CHh) * 8 + k16 % 8] = half_bits(v * up);
                            }
                double rest = (double)(a * L.bias[co]) * (double)ldexpf(1.f, ae - 16);
                for (int ky = 0; ky < 3; ++ky) {
                    const __half piece = __float2half_rn((float)rest);
                    rest -= (double)__half2float(piece);
                    for (int dx = 0; dx < 3; ++dx)
                        wq[(((size_t)(2 * ky + 1)) * 3 * C + (co / CHh) * 3 * CHh + dx * CHh + co % CHh) * 8 + 7] = __half_as_ushort(piece);
                }
                const uint16_t sc = half_bits((L.scale[co] < 0.f ? -1.f : 1.f) * ldexpf(1.f, 16 - ae)), sh = half_bits(L.shift[co]);
                uint32_t &ps = par[co / 2], &pt = par[C / 2 + co / 2];
                ps |= (uint32_t)sc << (16 * (co & 1));
                pt |= (uint32_t)sh << (16 * (co & 1));
                if (!std::isfinite(L.shift[co]) || fabsf(L.shift[co]) > 60000.f) ok = false;
            }
            if (ok) {
                if (int rc = upload_bytes(net, wq.data(), wq.size() * 2, &tc->d_w1_perm16)) return rc;
                if (int rc = upload_bytes(net, par.data(), par.size() * 4, reinterpret_cast<void **>(&tc->d_c1_par16))) return rc;
            }
        }
        {   // the taps as given (no BatchNorm scale, no bias row) for the batch-statistics forward pass
            std::vector<uint16_t> wp((size_t)6 * 3 * C * 8, 0);
            for (int co = 0; co < C; ++co)
                for (int ky = 0; ky < 3; ++ky)
                    for (int col = 0; col < 5; ++col)
                        for (int ch = 0; ch < 3; ++ch)
                            for (int dx = 0; dx < 3; ++dx) {
                                const int kx = col - dx, k16 = col * 3 + ch;
                                if (kx < 0 || kx > 2) continue;
                                wp[(((size_t)(2 * ky + k16 / 8)) * 3 * C + dx * C + co) * 8 + k16 % 8] =
                                    operand_bits(L.w[((size_t)co * 3 + ch) * 9 + ky * 3 + kx] * kW1Scale);
                            }
            if (int rc = upload_bytes(net, wp.data(), wp.size() * 2, &tc->d_w1_plain)) return rc;
            const std::vector<float> ones(C, 1.f), zeros(C, 0.f);
            if (int rc = upload_bytes(net, ones.data(), C * 4, reinterpret_cast<void **>(&tc->d_ones))) return rc;
            if (int rc = upload_bytes(net, zeros.data(), C * 4, reinterpret_cast<void **>(&tc->d_zeros))) return rc;
            if (int rc = upload_bytes(net, zeros.data(), C * 4, reinterpret_cast<void **>(&tc->d_bn_scale))) return rc;
            if (int rc = upload_bytes(net, zeros.data(), C * 4, reinterpret_cast<void **>(&tc->d_bn_shift))) return rc;
            const std::vector<double> dz(2 * C, 0.0);
            if (int rc = upload_bytes(net, dz.data(), dz.size() * 8, reinterpret_cast<void **>(&tc->d_bn_sums))) return rc;
        }
        if (int rc = upload_bytes(net, pb.data(), C * 4, reinterpret_cast<void **>(&tc->d_c1_bias))) return rc;
        if (int rc = upload_bytes(net, ps.data(), C * 4, reinterpret_cast<void **>(&tc->d_c1_sign))) return rc;
        if (int rc = upload_bytes(net, ps256.data(), C * 4, reinterpret_cast<void **>(&tc->d_c1_sign256))) return rc;
    }
    // conv2/3 B operand: [ky][ci/8][n' = blk*C + co, blk <-> kx = 2-blk][8]
    for (int layer = 1; layer <= 2; ++layer) {
        std::vector<uint16_t> w((size_t)3 * CG * 3 * C * 8, 0);
        const ConvLayer &L = net->conv[layer];
        for (int ky = 0; ky < 3; ++ky)
            for (int ci = 0; ci < C; ++ci)
                for (int blk = 0; blk < 3; ++blk)
                    for (int co = 0; co < C; ++co) {
                        const float v = L.w[((size_t)co * C + ci) * 9 + ky * 3 + (2 - blk)];
                        w[((((size_t)ky * CG + ci / 8) * 3 * C) + blk * C + co) * 8 + ci % 8] = operand_bits(v);
                    }
        if (int rc = upload_bytes(net, w.data(), w.size() * 2, layer == 1 ? &tc->d_w2 : &tc->d_w3)) return rc;
    }
    return C == 48 ? set_smem_limits<48>() : set_smem_limits<32>();
}

void tc_destroy(cutdet_net *net) {
    delete net->tc;
    net->tc = nullptr;
}

size_t tc_workspace_bytes(const cutdet_net *net, int batch, int height, int width) {
    const Geom g = make_geom(height, width, net->cfg.hidden_channels);
    return tc_workspace(net, g, batch).total + 1024;
}

int tc_forward_f32(cutdet_net *net, const float *x, int batch, int height, int width, float *logits, char *ws, cudaStream_t stream) {
    const Geom g = make_geom(height, width, net->cfg.hidden_channels);
    const TcWorkspace w = tc_workspace(net, g, batch);
    auto pack = [&](int f0, int nb) -> int {
        const int64_t total = (int64_t)nb * g.H * g.P1w;
        {
            KernelScope scope("pack_xin_f32", stream);
            pack_xin_f32_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(x + (size_t)f0 * 3 * height * width, nb, g.H, g.W,
                                                                                   g.P1w, reinterpret_cast<uint4 *>(ws + w.xin));
        }
        CUTDET_LAUNCH_CHECK("pack_xin_f32_kernel");
        return CUTDET_OK;
    };
    return g.C == 48 ? run_batch<48>(net, g, w, ws, batch, logits, stream, pack) : run_batch<32>(net, g, w, ws, batch, logits, stream, pack);
}

// Training-mode forward of the whole stack on the tensor-core path (one sub-batch: the statistics span the batch, so every
// frame's activations must be resident together).  Layer by layer: kernel with the identity affine, batch statistics, affine.
template <int C>
int run_batchstats(cutdet_net *net, const Geom &g, const TcWorkspace &w, char *ws, const float *x, int batch, float *out, cudaStream_t stream) {
    TcState *tc = net->tc;
    std::pair<CUtensorMap, CUtensorMap> *maps = nullptr;
    if (int rc = get_maps(net, g, w, ws, &maps)) return rc;
    auto phase_split_bn = [&](char *act, int gtot, int FP, int PW, int out_h, int out_w, const ConvLayer &L) -> int {
        const dim3 grid_s(std::min(64, (int)ceil_div(gtot, 256)), 9 * g.CG), grid_a((unsigned)ceil_div(gtot, 256), 9 * g.CG);
        {
            KernelScope scope("bn_ps_stats", stream);
            bn_ps_stats_kernel<<<grid_s, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(act), g.CG, gtot, batch * FP, tc->d_bn_sums);
        }
        CUTDET_LAUNCH_CHECK("bn_ps_stats_kernel");
        {
            KernelScope scope("bn_finalize", stream);
            bn_finalize_kernel<<<1, 64, 0, stream>>>(tc->d_bn_sums, (double)batch * out_h * out_w, L.d_gamma, L.d_beta, L.eps, C, tc->d_bn_scale,
                                                     tc->d_bn_shift);
        }
        CUTDET_LAUNCH_CHECK("bn_finalize_kernel");
        {
            KernelScope scope("bn_ps_apply", stream);
            bn_ps_apply_kernel<<<grid_a, 256, 0, stream>>>(reinterpret_cast<uint4 *>(act), g.CG, gtot, batch, FP, PW, out_h, out_w, tc->d_bn_scale,
                                                           tc->d_bn_shift);
        }
        CUTDET_LAUNCH_CHECK("bn_ps_apply_kernel");
        return CUTDET_OK;
    };
    // layer 1 (float input: the x-unfolded operand goes through L2)
    {
        const int64_t total = (int64_t)batch * g.H * g.P1w;
        KernelScope scope("pack_xin_f32", stream);
        pack_xin_f32_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(x, batch, g.H, g.W, g.P1w, reinterpret_cast<uint4 *>(ws + w.xin));
    }
    CUTDET_LAUNCH_CHECK("pack_xin_f32_kernel");
    Conv1Params c1;
    memset(&c1, 0, sizeof(c1));
    c1.xin = reinterpret_cast<const uint4 *>(ws + w.xin);
    c1.B = batch; c1.H = g.H; c1.P1h = g.P1h; c1.P1w = g.P1w;
    c1.tiles_per_frame = (g.P1h * g.P1w + 127) / 128;
    c1.out = OutSpec{ws + w.act1, 0, w.gtot1, g.PW1, g.FP1, g.Q1h, 0, g.P1h, g.P1w};
    c1.w_packed = reinterpret_cast<const uint4 *>(tc->d_w1_plain);
    c1.bias = net->conv[0].d_bias; c1.scale = tc->d_ones; c1.shift = tc->d_zeros;
    if (int rc = launch_conv1<C>(c1, stream)) return rc;
    if (int rc = phase_split_bn(ws + w.act1, w.gtot1, g.FP1, g.PW1, g.P1h, g.P1w, net->conv[0])) return rc;
    // layer 2
    MidParams p2 = mid_params(batch, g.FP1, g.PW1, g.P2h, g.P2w);
    p2.out = OutSpec{ws + w.act2, 0, w.gtot2, g.PW2, g.FP2, g.Q2h, 0, g.P2h, g.P2w};
    p2.w_packed = reinterpret_cast<const uint4 *>(tc->d_w2);
    p2.bias = net->conv[1].d_bias; p2.scale = tc->d_ones; p2.shift = tc->d_zeros;
    if (int rc = launch_mid<C>(maps->first, p2, "conv2_tc", stream, !net->opt.no_pdl)) return rc;
    if (int rc = phase_split_bn(ws + w.act2, w.gtot2, g.FP2, g.PW2, g.P2h, g.P2w, net->conv[1])) return rc;
    // layer 3: [frame][pixel][C] float32
    MidParams p3 = mid_params(batch, g.FP2, g.PW2, g.P3h, g.P3w);
    p3.out = OutSpec{ws + w.act3, 1, 0, 0, 0, 0, 0, g.P3h, g.P3w};
    p3.w_packed = reinterpret_cast<const uint4 *>(tc->d_w3);
    p3.bias = net->conv[2].d_bias; p3.scale = tc->d_ones; p3.shift = tc->d_zeros;
    if (int rc = launch_mid<C>(maps->second, p3, "conv3_tc", stream, !net->opt.no_pdl)) return rc;
    if (int rc = launch_bn_batchstats(reinterpret_cast<float *>(ws + w.act3), batch * g.P3h * g.P3w, C, 1, net->conv[2].d_gamma,
                                      net->conv[2].d_beta, net->conv[2].eps, stream))
        return rc;
    // head: avg-pool folded into the first matrix; every BatchNorm1d on batch statistics
    const float *cur = reinterpret_cast<const float *>(ws + w.act3);
    const int n_feat = g.P3h * g.P3w * g.C;
    float *folded = nullptr;
    if (int rc = folded_fc1(net, g, &folded)) return rc;
    if (net->cfg.n_fc_layers == 0) {
        KernelScope scope("head_fc1", stream);
        head_fc1_kernel<<<(batch + HEAD_FRAMES - 1) / HEAD_FRAMES, HEAD_THREADS, 0, stream>>>(
            cur, folded, tc->d_zero32, nullptr, nullptr, batch, n_feat, g.C * net->cfg.avg_pool_size * net->cfg.avg_pool_size, 0, out);
        CUTDET_LAUNCH_CHECK("head_fc1_kernel");
        return CUTDET_OK;
    }
    for (int j = 0; j < net->cfg.n_fc_layers; ++j) {
        const FcLayer &L = net->fc[j];
        const bool is_last = j + 1 == net->cfg.n_fc_layers;
        float *o = is_last ? out : reinterpret_cast<float *>(ws + w.fc[j & 1]);
        if (j == 0) {
            KernelScope scope("head_fc1", stream);
            head_fc1_kernel<<<(batch + HEAD_FRAMES - 1) / HEAD_FRAMES, HEAD_THREADS, 0, stream>>>(cur, folded, L.d_bias, nullptr, nullptr, batch, n_feat, L.out,
                                                                                        is_last ? 0 : 1, o);
            CUTDET_LAUNCH_CHECK("head_fc1_kernel");
        } else if (int rc = launch_fc(cur, o, L, batch, !is_last, stream, false)) {
            return rc;
        }
        if (L.has_bn)
            if (int rc = launch_bn_batchstats(o, batch, L.out, 1, L.d_gamma, L.d_beta, L.eps, stream)) return rc;
        cur = o;
    }
    return CUTDET_OK;
}

bool tc_batchstats_supported(const cutdet_net *net, int batch, int height, int width) {
    // everything resident in one pass: the workspace (tc_workspace) holds `sub` frames of input and layer-1 activations
    const int sub = net->opt.sub_batch > 0 ? net->opt.sub_batch : SUB_BATCH;
    return tc_supported(net, height, width) && batch <= BATCHSTATS_MAX && batch <= sub;
}

int tc_forward_f32_batchstats(cutdet_net *net, const float *x, int batch, int height, int width, float *out, char *ws, cudaStream_t stream) {
    const Geom g = make_geom(height, width, net->cfg.hidden_channels);
    const TcWorkspace w = tc_workspace(net, g, batch);
    return g.C == 48 ? run_batchstats<48>(net, g, w, ws, x, batch, out, stream) : run_batchstats<32>(net, g, w, ws, x, batch, out, stream);
}

int tc_forward_frames(cutdet_net *net, const cutdet_resize_plan *plan, const cutdet_frames *src, float *logits, char *ws,
                      cudaStream_t stream) {
    const int batch = src->batch;
    const Geom g = make_geom(plan->host.dst_h, plan->host.dst_w, net->cfg.hidden_channels);
    const TcWorkspace w = tc_workspace(net, g, batch);
    auto pack = [&](int f0, int nb) -> int {
        const int64_t total = (int64_t)nb * g.H * g.P1w;
        {
            KernelScope scope("preprocess_xin", stream);
            preprocess_xin_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(
                plan->host, src->frames_dev + (int64_t)f0 * src->frame_stride, src->frame_stride, src->row_pitch,
                src->row_map_compact, nb, g.P1w, reinterpret_cast<uint4 *>(ws + w.xin));
        }
        CUTDET_LAUNCH_CHECK("preprocess_xin_kernel");
        return CUTDET_OK;
    };
    FusedSrc fs;
    const FusedSrc *fused = (net->tc->c1_folded && fused_source(plan, src, g, &fs)) ? &fs : nullptr;   // the fused kernel needs the bias row
    return g.C == 48 ? run_batch<48>(net, g, w, ws, batch, logits, stream, pack, fused)
                     : run_batch<32>(net, g, w, ws, batch, logits, stream, pack, fused);
}

int tc_debug_conv_output(cutdet_net *net, int layer, int batch, int height, int width, const char *ws, float *out,
                         cudaStream_t stream) {
    const Geom g = make_geom(height, width, net->cfg.hidden_channels);
    const TcWorkspace w = tc_workspace(net, g, batch);
    if (batch > w.sub && layer < 2)
        return fail(CUTDET_EUNSUPPORTED, "debug_conv_output: layers 0 and 1 are only kept for batches of up to %d frames", w.sub);
    if (layer == 2) {
        const int64_t total = (int64_t)batch * g.C * g.P3h * g.P3w;
        unpack_plain_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(reinterpret_cast<const float *>(ws + w.act3), batch, g.C,
                                                                               g.P3h * g.P3w, out);
    } else {
        const int ph = layer == 0 ? g.P1h : g.P2h, pw = layer == 0 ? g.P1w : g.P2w;
        const int64_t total = (int64_t)batch * g.C * ph * pw;
        unpack_phase_split_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(
            reinterpret_cast<const uint16_t *>(ws + (layer == 0 ? w.act1 : w.act2)), layer == 0 ? w.gtot1 : w.gtot2,
            layer == 0 ? g.FP1 : g.FP2, layer == 0 ? g.PW1 : g.PW2, batch, g.C, ph, pw, out);
    }
    CUTDET_LAUNCH_CHECK("unpack kernel");
    return CUTDET_OK;
}

}  // namespace cutdet
