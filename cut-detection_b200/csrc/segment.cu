// K4 (per-frame max/argmax), K5 (streaming run-length encoding), K6 (orphan gluing, adjacent merge) and the
// shard stitch.  Stands in for reference frameID/segmentation.py:
//   K4  torch.max(scores, dim=1)                                     :37
//   K5  run boundaries, run lengths, per-run mean of the max logit   :39-60
//   K6  glue_orphans (:91-166, _find_orphans :12-17, _update_neighbor :69-89), combine_adjacent_segments (:168-183)
// These are integer/byte passes over 5 bytes per frame: HBM/latency bound, no tensor-core work.
#include <algorithm>

#include "common.cuh"

namespace cutdet {

namespace {

constexpr int RLE_THREADS = 256;
constexpr int RLE_ITEMS = 8;
constexpr int RLE_TILE = RLE_THREADS * RLE_ITEMS;   // frames per block
constexpr int RLE_MAX_BLOCKS = 1024;                // frames per launch <= 2,097,152 (host loops beyond that)

// What a prefix of the frame sequence leaves behind: closed-run count and the still-open run.
struct Prefix {
    long long count;   // runs closed so far
    long long start;   // first frame of the open run
    double sum;        // sum of max logits over the open run
    int has;           // (scan element only) a run boundary lies inside this element
};

__device__ __forceinline__ Prefix combine(const Prefix &a, const Prefix &b) {   // a first, then b
    Prefix r;
    r.count = a.count + b.count;
    r.start = b.has ? b.start : a.start;
    r.sum = b.has ? b.sum : a.sum + b.sum;
    r.has = a.has | b.has;
    return r;
}

__device__ __forceinline__ Prefix shfl_up(const Prefix &p, int d) {
    Prefix r;
    r.count = __shfl_up_sync(0xffffffffu, p.count, d);
    r.start = __shfl_up_sync(0xffffffffu, p.start, d);
    r.sum = __shfl_up_sync(0xffffffffu, p.sum, d);
    r.has = __shfl_up_sync(0xffffffffu, p.has, d);
    return r;
}

struct BlockSlot {
    Prefix inclusive;
    volatile int ready;
    int pad;
};

struct RleState {
    long long n_frames;
    long long n_closed;
    long long open_start;
    double open_sum;
    int open_type;      // label of the open run; -1 before the first frame
    int overflow;
    unsigned int ticket;
    unsigned int pad;
    BlockSlot slots[RLE_MAX_BLOCKS];
};

// ------------------------------------------------------------------------------------------- K4
template <int C>
__global__ void argmax_kernel(const float *__restrict__ scores, int64_t n, int n_classes, uint8_t *__restrict__ labels,
                              float *__restrict__ top) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nc = C > 0 ? C : n_classes;
    const float *s = scores + i * nc;
    float best = s[0];
    int arg = 0;
#pragma unroll
    for (int c = 1; c < nc; ++c) {
        const float v = s[c];
        if (v > best) { best = v; arg = c; }     // strict: the first maximum wins, as torch.max on CPU
    }
    labels[i] = (uint8_t)arg;
    top[i] = best;
}

// ------------------------------------------------------------------------------------------- K5
// One launch appends n frames.  Blocks take tiles in ticket order; each computes its local aggregate, waits for
// its predecessor's inclusive prefix (chained scan: a block only ever waits on blocks that started earlier),
// publishes its own, then writes the runs that close inside its tile.
__global__ void __launch_bounds__(RLE_THREADS)
rle_append_kernel(RleState *st, const uint8_t *__restrict__ labels, const float *__restrict__ top, long long n,
                  cutdet_run_table table) {
    __shared__ unsigned int s_block;
    __shared__ Prefix s_warp[RLE_THREADS / 32];
    __shared__ Prefix s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_block = atomicAdd(&st->ticket, 1u);
    __syncthreads();
    const unsigned int block = s_block;
    const unsigned int n_blocks = gridDim.x;
    // state as it was BEFORE this launch (only the last block rewrites it, after every block has read it)
    const long long frames_before = st->n_frames;
    const int type_before = st->open_type;

    const long long i0 = (long long)block * RLE_TILE + (long long)tid * RLE_ITEMS;
    int lab[RLE_ITEMS + 1];                       // lab[0] = label of the frame before my first item
    float val[RLE_ITEMS];
    int n_mine = 0;
    if (i0 < n) {
        n_mine = (int)min((long long)RLE_ITEMS, n - i0);
        lab[0] = i0 > 0 ? (int)labels[i0 - 1] : type_before;
#pragma unroll
        for (int k = 0; k < RLE_ITEMS; ++k) {
            if (k < n_mine) { lab[k + 1] = labels[i0 + k]; val[k] = top[i0 + k]; }
            else { lab[k + 1] = -1; val[k] = 0.f; }
        }
    }
    // local element: boundaries inside my items
    Prefix mine;
    mine.count = 0; mine.start = 0; mine.sum = 0.0; mine.has = 0;
    unsigned int bmask = 0;
#pragma unroll
    for (int k = 0; k < RLE_ITEMS; ++k) {
        if (k < n_mine) {
            const bool first_ever = (frames_before == 0 && i0 + k == 0);
            const bool boundary = !first_ever && lab[k + 1] != lab[k];
            if (boundary) {
                bmask |= 1u << k;
                mine.count += 1;
                mine.has = 1;
                mine.start = frames_before + i0 + k;
                mine.sum = 0.0;
            }
            mine.sum += (double)val[k];
        }
    }
    // block-wide inclusive scan of the elements
    Prefix inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Prefix o = shfl_up(inc, d);
        if (lane >= d) inc = combine(o, inc);
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        Prefix w;
        if (lane < RLE_THREADS / 32) w = s_warp[lane];
        else { w.count = 0; w.start = 0; w.sum = 0.0; w.has = 0; }
#pragma unroll
        for (int d = 1; d < RLE_THREADS / 32; d <<= 1) {
            Prefix o = shfl_up(w, d);
            if (lane >= d) w = combine(o, w);
        }
        if (lane < RLE_THREADS / 32) s_warp[lane] = w;     // inclusive over warps
    }
    __syncthreads();
    // predecessor's inclusive prefix
    if (tid == 0) {
        Prefix carry;
        if (block == 0) {
            carry.count = st->n_closed; carry.start = st->open_start; carry.sum = st->open_sum; carry.has = 0;
        } else {
            BlockSlot *prev = &st->slots[block - 1];
            while (prev->ready == 0) { __nanosleep(20); }
            __threadfence();
            // L2 loads: another tile on this SM may have pulled the same line into L1 before it was complete
            carry.count = __ldcg(&prev->inclusive.count);
            carry.start = __ldcg(&prev->inclusive.start);
            carry.sum = __ldcg(&prev->inclusive.sum);
            carry.has = 0;
        }
        s_carry = carry;
        Prefix total = combine(carry, s_warp[RLE_THREADS / 32 - 1]);
        total.has = 0;
        if (block + 1 < n_blocks) {
            st->slots[block].inclusive = total;
            __threadfence();
            st->slots[block].ready = 1;
        } else {
            // last tile of the launch: persist the stream state, re-arm the scan workspace
            st->n_closed = total.count;
            st->open_start = total.start;
            st->open_sum = total.sum;
            st->open_type = (int)labels[n - 1];
            st->n_frames = frames_before + n;
            for (unsigned int b = 0; b + 1 < n_blocks; ++b) st->slots[b].ready = 0;
            st->ticket = 0;
            __threadfence();
        }
    }
    __syncthreads();
    // my exclusive prefix = carry (+) warps before mine (+) lanes before me
    Prefix excl = s_carry;
    if (wid > 0) excl = combine(excl, s_warp[wid - 1]);
    {
        Prefix o = shfl_up(inc, 1);
        if (lane > 0) excl = combine(excl, o);
    }
    if (n_mine == 0 || bmask == 0) return;
    long long idx = excl.count;
    long long start = excl.start;
    double sum = excl.sum;
#pragma unroll
    for (int k = 0; k < RLE_ITEMS; ++k) {
        if (k < n_mine) {
            if (bmask & (1u << k)) {
                const long long end = frames_before + i0 + k - 1;
                if (idx < table.capacity) {
                    const long long len = end - start + 1;
                    table.end_frames_dev[idx] = end;
                    table.start_frames_dev[idx] = start;
                    table.run_lengths_dev[idx] = len;
                    table.frame_types_dev[idx] = lab[k];
                    table.score_sums_dev[idx] = sum;
                    table.score_means_dev[idx] = (float)(sum / (double)len);
                } else {
                    st->overflow = 1;
                }
                idx += 1;
                start = end + 1;
                sum = 0.0;
            }
            sum += (double)val[k];
        }
    }
}

__global__ void rle_finish_kernel(RleState *st, cutdet_run_table table, int64_t *n_runs) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long count = st->n_closed;
    if (st->n_frames > 0) {
        if (count < table.capacity) {
            const long long end = st->n_frames - 1, len = end - st->open_start + 1;
            table.end_frames_dev[count] = end;
            table.start_frames_dev[count] = st->open_start;
            table.run_lengths_dev[count] = len;
            table.frame_types_dev[count] = st->open_type;
            table.score_sums_dev[count] = st->open_sum;
            table.score_means_dev[count] = (float)(st->open_sum / (double)len);
        } else {
            st->overflow = 1;
        }
        count += 1;
    }
    if (n_runs) *n_runs = count;
}

// ------------------------------------------------------------------------------------------- K6
// Orphan gluing is sequential by construction (least confident orphan first, and every merge changes the queue),
// so it runs as ONE warp: the warp keeps a 32-ary tournament tree of the orphans' means (level 0: one key per run,
// +inf when the run is dead or not an orphan; level l: min of 32 children with the run index that attains it) and
// each step is root lookup -> merge by lane 0 -> two leaf-to-root refreshes, each level one coalesced 32-wide load
// and a shuffle reduction.  Ties go to the lowest run index.
// Keys are the means mapped to unsigned integers of the same order (so a NaN mean can be given a place: LAST among the orphans,
// where torch.argsort puts it -- the reference still glues such a run, after every orphan with a number for a mean), KEY_NONE
// = not an orphan / dead.
struct GlueScratch {
    int *prev, *next;        // neighbours in the list of live runs; prev == -2 marks a dead run
    uint32_t *key0;          // level 0 keys
    uint32_t *lv_val[4];     // levels 1..4 (sized for the table's capacity)
    int *lv_idx[4];
    long long lv_n[5];       // entries per level for the CURRENT run count (lv_n[0] = S), set by the kernel
    int n_levels;            // levels above 0 in use, set by the kernel
};

__device__ __forceinline__ void glue_levels(GlueScratch &ws, long long S) {
    ws.lv_n[0] = S;
    int levels = 0;
    while (ws.lv_n[levels] > 32 && levels < 4) { ws.lv_n[levels + 1] = (ws.lv_n[levels] + 31) / 32; ++levels; }
    ws.n_levels = levels;
}

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu, KEY_NAN = 0xFFFFFFFEu;

__device__ __forceinline__ uint32_t orphan_key(float mean) {
    if (mean != mean) return KEY_NAN;
    const uint32_t u = __float_as_uint(mean + 0.0f);           // -0 -> +0: equal floats keep equal keys
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ bool key_less(uint32_t va, int ia, uint32_t vb, int ib) { return va < vb || (va == vb && ia < ib); }

__device__ __forceinline__ void warp_argmin(uint32_t &v, int &i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const uint32_t ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int oi = __shfl_xor_sync(0xffffffffu, i, off);
        if (key_less(ov, oi, v, i)) { v = ov; i = oi; }
    }
}

__device__ __forceinline__ bool is_orphan(int type, long long len, int k_real, int k_blank) {
    return type != 2 ? len < k_real : len < k_blank;        // segmentation.py:12-17
}

// Recompute the ancestors of leaf `leaf` (all lanes participate).
__device__ void refresh_path(const GlueScratch &ws, long long leaf, int lane) {
    long long child = leaf;
    for (int l = 0; l < ws.n_levels; ++l) {
        const long long group = child >> 5;
        const long long c = (group << 5) + lane;
        uint32_t v = KEY_NONE;
        int i = 0x7fffffff;
        if (c < ws.lv_n[l]) {
            if (l == 0) { v = ((volatile uint32_t *)ws.key0)[c]; i = (int)c; }
            else { v = ((volatile uint32_t *)ws.lv_val[l - 1])[c]; i = ((volatile int *)ws.lv_idx[l - 1])[c]; }
        }
        warp_argmin(v, i);
        if (lane == 0) {
            ((volatile uint32_t *)ws.lv_val[l])[group] = v;
            ((volatile int *)ws.lv_idx[l])[group] = i;
        }
        __syncwarp();
        child = group;
    }
}

__global__ void __launch_bounds__(32)
glue_orphans_kernel(cutdet_run_table t, int64_t *n_runs, int k_real, int k_blank, int32_t *status, GlueScratch ws) {
    const int lane = threadIdx.x;
    const long long S = *n_runs;
    if (lane == 0) *status = CUTDET_OK;
    if (S <= 0) return;
    glue_levels(ws, S);      // the tree covers the S live rows, not the table's capacity
    volatile int64_t *start = t.start_frames_dev, *end = t.end_frames_dev, *len = t.run_lengths_dev;
    volatile int32_t *type = t.frame_types_dev;
    volatile float *mean = t.score_means_dev;
    volatile int *prev = ws.prev, *next = ws.next;
    volatile uint32_t *key0 = ws.key0;
    // build
    for (long long i = lane; i < S; i += 32) {
        prev[i] = (int)i - 1;
        next[i] = i + 1 < S ? (int)i + 1 : -1;
        key0[i] = is_orphan(type[i], len[i], k_real, k_blank) ? orphan_key(mean[i]) : KEY_NONE;
    }
    __syncwarp();
    for (int l = 0; l < ws.n_levels; ++l) {
        for (long long g = 0; g < ws.lv_n[l + 1]; ++g) {
            const long long c = (g << 5) + lane;
            uint32_t v = KEY_NONE;
            int i = 0x7fffffff;
            if (c < ws.lv_n[l]) {
                if (l == 0) { v = key0[c]; i = (int)c; }
                else { v = ((volatile uint32_t *)ws.lv_val[l - 1])[c]; i = ((volatile int *)ws.lv_idx[l - 1])[c]; }
            }
            warp_argmin(v, i);
            if (lane == 0) { ((volatile uint32_t *)ws.lv_val[l])[g] = v; ((volatile int *)ws.lv_idx[l])[g] = i; }
        }
        __syncwarp();
    }
    // glue
    bool lone = false;
    while (true) {
        uint32_t v = KEY_NONE;
        int target = 0x7fffffff;
        const int top = ws.n_levels;     // top level has <= 32 entries
        if (lane < ws.lv_n[top]) {
            if (top == 0) { v = key0[lane]; target = lane; }
            else { v = ((volatile uint32_t *)ws.lv_val[top - 1])[lane]; target = ((volatile int *)ws.lv_idx[top - 1])[lane]; }
        }
        warp_argmin(v, target);
        if (v == KEY_NONE) break;                   // no orphan left
        int nb = -1;
        if (lane == 0) {
            const int p = prev[target], n = next[target];
            if (p < 0 && n < 0) {
                nb = -3;                            // a lone orphan: the reference raises IndexError here
            } else {
                if (p < 0) nb = n;
                else if (n < 0) nb = p;
                else nb = len[p] > len[n] ? p : n;  // the longer neighbour; ties go to the next run
                if (target < nb) start[nb] = start[target]; else end[nb] = end[target];
                const float ln = (float)len[nb], lo = (float)len[target];
                const float num = __fadd_rn(__fmul_rn(mean[nb], ln), __fmul_rn(mean[target], lo));
                mean[nb] = __fadd_rn(__fdiv_rn(num, ln), lo);           // (m_n*l_n + m_o*l_o) / l_n + l_o, as written
                len[nb] = end[nb] - start[nb] + 1;
                if (p >= 0) next[p] = n;
                if (n >= 0) prev[n] = p;
                prev[target] = -2;
                key0[target] = KEY_NONE;
                key0[nb] = is_orphan(type[nb], len[nb], k_real, k_blank) ? orphan_key(mean[nb]) : KEY_NONE;
            }
        }
        nb = __shfl_sync(0xffffffffu, nb, 0);
        if (nb == -3) { lone = true; break; }
        __syncwarp();
        refresh_path(ws, target, lane);
        // a neighbour in the same level-0 group was already covered by the refresh above
        if ((nb >> 5) != (target >> 5)) refresh_path(ws, nb, lane);
    }
    if (lone) {
        if (lane == 0) *status = CUTDET_ELONE_ORPHAN;
        return;
    }
    // compact the live runs, in order
    long long out = 0;
    for (long long base = 0; base < S; base += 32) {
        const long long i = base + lane;
        const bool live = i < S && prev[i] != -2;
        int64_t s = 0, e = 0, l = 0; int32_t ty = 0; float m = 0.f;
        if (live) { s = start[i]; e = end[i]; l = len[i]; ty = type[i]; m = mean[i]; }
        const unsigned int ballot = __ballot_sync(0xffffffffu, live);
        __syncwarp();
        if (live) {
            const long long o = out + __popc(ballot & ((1u << lane) - 1));
            start[o] = s; end[o] = e; len[o] = l; type[o] = ty; mean[o] = m;
        }
        out += __popc(ballot);
        __syncwarp();
    }
    if (lane == 0) *n_runs = out;
}

// Adjacent runs of equal type: the reference merges the first matching pair (i, i+1) into i+1 and repeats, i.e. it
// folds each maximal group left to right.  One lane per group end does that fold; groups are independent.
__global__ void __launch_bounds__(32) combine_adjacent_kernel(cutdet_run_table t, int64_t *n_runs) {
    const int lane = threadIdx.x;
    const long long S = *n_runs;
    if (S <= 0) return;
    volatile int64_t *start = t.start_frames_dev, *end = t.end_frames_dev, *len = t.run_lengths_dev;
    volatile int32_t *type = t.frame_types_dev;
    volatile float *mean = t.score_means_dev;
    long long out = 0;
    for (long long base = 0; base < S; base += 32) {
        const long long i = base + lane;
        bool is_end = false;
        int64_t s = 0, e = 0, l = 0; int32_t ty = 0; float m = 0.f;
        if (i < S) {
            ty = type[i];
            is_end = (i + 1 == S) || (type[i + 1] != ty);
            if (is_end) {
                long long j = i;
                while (j > 0 && type[j - 1] == ty) --j;
                s = start[j]; m = mean[j]; l = len[j];
                for (long long k = j + 1; k <= i; ++k) {
                    const float ln = (float)len[k], lo = (float)l;
                    const float num = __fadd_rn(__fmul_rn(mean[k], ln), __fmul_rn(m, lo));
                    m = __fadd_rn(__fdiv_rn(num, ln), lo);
                    l = end[k] - s + 1;
                }
                e = end[i];
            }
        }
        const unsigned int ballot = __ballot_sync(0xffffffffu, is_end);
        __syncwarp();
        if (is_end) {
            const long long o = out + __popc(ballot & ((1u << lane) - 1));
            start[o] = s; end[o] = e; len[o] = l; type[o] = ty; mean[o] = m;
        }
        out += __popc(ballot);
        __syncwarp();
    }
    if (lane == 0) *n_runs = out;
}

// ------------------------------------------------------------------------------------------- shard stitch
__global__ void __launch_bounds__(256)
stitch_kernel(cutdet_run_table src, int n_shards, int64_t shard_capacity, const int64_t *n_runs, const int64_t *offsets,
              cutdet_run_table dst, int64_t *n_out) {
    __shared__ long long s_out;
    if (threadIdx.x == 0) s_out = 0;
    __syncthreads();
    for (int sh = 0; sh < n_shards; ++sh) {
        const long long n = n_runs[sh];
        if (n <= 0) continue;                      // uniform across the block
        const long long base = (long long)sh * shard_capacity, off = offsets[sh];
        const long long out0 = s_out;
        const bool join = out0 > 0 && dst.frame_types_dev[out0 - 1] == src.frame_types_dev[base];
        __syncthreads();
        const long long skip = join ? 1 : 0;
        if (threadIdx.x == 0 && join) {
            const long long o = out0 - 1;
            const double sum = dst.score_sums_dev[o] + src.score_sums_dev[base];
            const long long e = off + src.end_frames_dev[base], len = e - dst.start_frames_dev[o] + 1;
            dst.end_frames_dev[o] = e;
            dst.run_lengths_dev[o] = len;
            dst.score_sums_dev[o] = sum;
            dst.score_means_dev[o] = (float)(sum / (double)len);
        }
        for (long long k = skip + threadIdx.x; k < n; k += blockDim.x) {
            const long long o = out0 + k - skip;
            if (o < dst.capacity) {
                dst.end_frames_dev[o] = off + src.end_frames_dev[base + k];
                dst.start_frames_dev[o] = off + src.start_frames_dev[base + k];
                dst.run_lengths_dev[o] = src.run_lengths_dev[base + k];
                dst.frame_types_dev[o] = src.frame_types_dev[base + k];
                dst.score_sums_dev[o] = src.score_sums_dev[base + k];
                dst.score_means_dev[o] = src.score_means_dev[base + k];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_out = out0 + n - skip;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = s_out;     // > dst.capacity means rows were dropped; the caller checks
}

// ------------------------------------------------------------------------------------------- packed shard tables
// What travels in the one exchange step of the multi-GPU path: a 16-byte header (n_runs, n_frames) and `capacity`
// rows of 40 bytes -- end, start, length (int64), sum (float64), type (int32), mean (float32) -- per shard.  Sums and lengths,
// not means, so a run cut by a shard edge is re-joined exactly.
struct __align__(8) PackedRow { long long end, start, length; double sum; int type; float mean; };
static_assert(sizeof(PackedRow) == 40, "packed run-table row");
constexpr size_t PACK_HEADER = 16;

__global__ void __launch_bounds__(256) shard_pack_kernel(cutdet_run_table t, const int64_t *n_runs, long long n_frames,
                                                         long long capacity, uint8_t *packed) {
    const long long n = *n_runs;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        reinterpret_cast<long long *>(packed)[0] = n;          // may exceed `capacity`: the stitch reports it
        reinterpret_cast<long long *>(packed)[1] = n_frames;
    }
    PackedRow *rows = reinterpret_cast<PackedRow *>(packed + PACK_HEADER);
    const long long m = n < capacity ? n : capacity;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        PackedRow r;
        r.end = t.end_frames_dev[i]; r.start = t.start_frames_dev[i]; r.length = t.run_lengths_dev[i];
        r.sum = t.score_sums_dev[i]; r.type = t.frame_types_dev[i]; r.mean = t.score_means_dev[i];
        rows[i] = r;
    }
}

// Joins the gathered shards (one block; S is tens to thousands).  Frame offsets are the running sum of the headers' frame
// counts.  *n_out < 0 reports a shard whose table did not fit the gather capacity: -(1 + its run count).
__global__ void __launch_bounds__(256)
stitch_packed_kernel(const uint8_t *__restrict__ gathered, int n_shards, long long capacity, cutdet_run_table dst, int64_t *n_out,
                     int64_t *total_frames) {
    __shared__ long long s_out;
    const size_t stride = PACK_HEADER + (size_t)capacity * sizeof(PackedRow);
    if (threadIdx.x == 0) s_out = 0;
    __syncthreads();
    long long off = 0, worst = 0;
    for (int sh = 0; sh < n_shards; ++sh) {
        const uint8_t *base = gathered + (size_t)sh * stride;
        const long long n = reinterpret_cast<const long long *>(base)[0], frames = reinterpret_cast<const long long *>(base)[1];
        const PackedRow *rows = reinterpret_cast<const PackedRow *>(base + PACK_HEADER);
        if (n > capacity && n > worst) worst = n;
        if (n > 0 && n <= capacity) {                      // uniform across the block
            const long long out0 = s_out;
            const bool join = out0 > 0 && dst.frame_types_dev[out0 - 1] == rows[0].type;
            __syncthreads();
            const long long skip = join ? 1 : 0;
            if (threadIdx.x == 0 && join) {
                const long long o = out0 - 1;
                const double sum = dst.score_sums_dev[o] + rows[0].sum;
                const long long e = off + rows[0].end, len = e - dst.start_frames_dev[o] + 1;
                dst.end_frames_dev[o] = e;
                dst.run_lengths_dev[o] = len;
                dst.score_sums_dev[o] = sum;
                dst.score_means_dev[o] = (float)(sum / (double)len);
            }
            for (long long k = skip + threadIdx.x; k < n; k += blockDim.x) {
                const long long o = out0 + k - skip;
                if (o < dst.capacity) {
                    const PackedRow r = rows[k];
                    dst.end_frames_dev[o] = off + r.end;
                    dst.start_frames_dev[o] = off + r.start;
                    dst.run_lengths_dev[o] = r.length;
                    dst.frame_types_dev[o] = r.type;
                    dst.score_sums_dev[o] = r.sum;
                    dst.score_means_dev[o] = r.mean;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_out = out0 + n - skip;
            __syncthreads();
        }
        off += frames;
    }
    if (threadIdx.x == 0) {
        *n_out = worst > 0 ? -(1 + worst) : s_out;        // > dst.capacity means rows were dropped; the caller checks
        if (total_frames) *total_frames = off;
    }
}

// ------------------------------------------------------------------------------------------- scratch for K6
// The caller owns it (cutdet_glue_orphans_workspace_bytes): prev[S] next[S] key0[S], then (key, index) per tree level.
int glue_level_sizes(long long capacity, long long (&n)[5]) {      // returns the levels above 0, -1 if the table is too large
    n[0] = capacity;
    int levels = 0;
    while (n[levels] > 32 && levels < 4) { n[levels + 1] = (n[levels] + 31) / 32; ++levels; }
    for (int l = levels + 1; l < 5; ++l) n[l] = 0;
    return n[levels] > 32 ? -1 : levels;
}

size_t glue_scratch_bytes(long long capacity) {
    long long n[5];
    const int levels = glue_level_sizes(capacity, n);
    if (levels < 0) return 0;
    size_t bytes = (size_t)capacity * 12;
    for (int l = 1; l <= levels; ++l) bytes += (size_t)n[l] * 8;
    return bytes + 256;
}

int glue_scratch(long long capacity, void *workspace, size_t workspace_bytes, GlueScratch *ws) {
    long long n[5];
    const int levels = glue_level_sizes(capacity, n);
    if (levels < 0) return fail(CUTDET_ECAPACITY, "glue_orphans: run table of %lld rows is too large", capacity);
    if (!workspace || workspace_bytes < glue_scratch_bytes(capacity))
        return fail(CUTDET_ECAPACITY, "glue_orphans: workspace of %zu bytes, %zu needed (cutdet_glue_orphans_workspace_bytes)",
                    workspace_bytes, glue_scratch_bytes(capacity));
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return fail(CUTDET_EINVAL, "glue_orphans: workspace must be 16-byte aligned");
    char *p = reinterpret_cast<char *>(workspace);
    ws->prev = reinterpret_cast<int *>(p); p += (size_t)capacity * 4;
    ws->next = reinterpret_cast<int *>(p); p += (size_t)capacity * 4;
    ws->key0 = reinterpret_cast<uint32_t *>(p); p += (size_t)capacity * 4;
    for (int l = 1; l <= levels; ++l) {
        ws->lv_val[l - 1] = reinterpret_cast<uint32_t *>(p); p += (size_t)n[l] * 4;
        ws->lv_idx[l - 1] = reinterpret_cast<int *>(p); p += (size_t)n[l] * 4;
    }
    for (int l = levels; l < 4; ++l) { ws->lv_val[l] = nullptr; ws->lv_idx[l] = nullptr; }
    for (int l = 0; l < 5; ++l) ws->lv_n[l] = n[l];
    ws->n_levels = levels;
    return CUTDET_OK;
}

int check_table(const cutdet_run_table *t, const char *who) {
    CUTDET_REQUIRE(t && t->end_frames_dev && t->start_frames_dev && t->run_lengths_dev && t->frame_types_dev &&
                       t->score_means_dev && t->score_sums_dev && t->capacity > 0,
                   "%s: incomplete run table", who);
    return CUTDET_OK;
}

}  // namespace
}  // namespace cutdet

using namespace cutdet;

extern "C" int cutdet_argmax(const float *scores, int64_t n, int n_classes, uint8_t *labels, float *top,
                             cutdet_stream_t stream) {
    CUTDET_REQUIRE(n >= 0 && n_classes >= 1 && n_classes <= 256, "argmax: bad shape [%lld, %d]", (long long)n, n_classes);
    if (n == 0) return CUTDET_OK;
    CUTDET_REQUIRE(scores && labels && top, "argmax: null pointer");
    const unsigned grid = (unsigned)ceil_div(n, 256);
    {
        KernelScope scope("argmax_kernel", as_stream(stream));
        if (n_classes == 3) argmax_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(scores, n, 3, labels, top);
        else argmax_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(scores, n, n_classes, labels, top);
    }
    CUTDET_LAUNCH_CHECK("argmax_kernel");
    return CUTDET_OK;
}

extern "C" size_t cutdet_rle_state_bytes(void) { return sizeof(RleState); }

extern "C" int cutdet_rle_reset(void *state, cutdet_stream_t stream) {
    CUTDET_REQUIRE(state, "rle_reset: null state");
    CUTDET_CUDA(cudaMemsetAsync(state, 0, sizeof(RleState), as_stream(stream)));
    // open_type = -1: no frame seen yet
    static const int minus_one = -1;
    CUTDET_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(state) + offsetof(RleState, open_type), &minus_one, sizeof(int),
                                cudaMemcpyHostToDevice, as_stream(stream)));
    return CUTDET_OK;
}

extern "C" int cutdet_rle_append(void *state, const uint8_t *labels, const float *top, int64_t n,
                                 const cutdet_run_table *table, cutdet_stream_t stream) {
    CUTDET_REQUIRE(state && n >= 0, "rle_append: bad argument");
    if (int rc = check_table(table, "rle_append")) return rc;
    if (n == 0) return CUTDET_OK;
    CUTDET_REQUIRE(labels && top, "rle_append: null labels/top");
    const int64_t per_launch = (int64_t)RLE_MAX_BLOCKS * RLE_TILE;
    for (int64_t off = 0; off < n; off += per_launch) {
        const int64_t m = n - off < per_launch ? n - off : per_launch;
        {
            KernelScope scope("rle_append_kernel", as_stream(stream));
            rle_append_kernel<<<(unsigned)ceil_div(m, RLE_TILE), RLE_THREADS, 0, as_stream(stream)>>>(
            reinterpret_cast<RleState *>(state), labels + off, top + off, (long long)m, *table);
        }
        CUTDET_LAUNCH_CHECK("rle_append_kernel");
    }
    return CUTDET_OK;
}

extern "C" int cutdet_rle_finish(void *state, const cutdet_run_table *table, int64_t *n_runs, cutdet_stream_t stream) {
    CUTDET_REQUIRE(state, "rle_finish: null state");
    if (int rc = check_table(table, "rle_finish")) return rc;
    {
        KernelScope scope("rle_finish_kernel", as_stream(stream));
        rle_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<RleState *>(state), *table, n_runs);
    }
    CUTDET_LAUNCH_CHECK("rle_finish_kernel");
    return CUTDET_OK;
}

extern "C" int cutdet_rle_count(const void *state, int64_t *n_runs_host, cutdet_stream_t stream) {
    CUTDET_REQUIRE(state && n_runs_host, "rle_count: null argument");
    struct { long long n_frames, n_closed, open_start; double open_sum; int open_type, overflow; } head;
    CUTDET_CUDA(cudaMemcpyAsync(&head, state, sizeof(head), cudaMemcpyDeviceToHost, as_stream(stream)));
    CUTDET_CUDA(cudaStreamSynchronize(as_stream(stream)));
    *n_runs_host = head.n_closed + (head.n_frames > 0 ? 1 : 0);
    if (head.overflow) return fail(CUTDET_ECAPACITY, "run table overflow: more than its capacity of runs");
    return CUTDET_OK;
}

extern "C" size_t cutdet_glue_orphans_workspace_bytes(int64_t capacity) { return capacity > 0 ? glue_scratch_bytes(capacity) : 0; }

extern "C" int cutdet_glue_orphans(const cutdet_run_table *table, int64_t *n_runs, int real_threshold, int blank_threshold,
                                   int32_t *status, void *workspace, size_t workspace_bytes, cutdet_stream_t stream) {
    if (int rc = check_table(table, "glue_orphans")) return rc;
    CUTDET_REQUIRE(n_runs && status, "glue_orphans: null n_runs/status");
    GlueScratch ws;
    if (int rc = glue_scratch(table->capacity, workspace, workspace_bytes, &ws)) return rc;
    {
        KernelScope scope("glue_orphans_kernel", as_stream(stream));
        glue_orphans_kernel<<<1, 32, 0, as_stream(stream)>>>(*table, n_runs, real_threshold, blank_threshold, status, ws);
    }
    CUTDET_LAUNCH_CHECK("glue_orphans_kernel");
    return CUTDET_OK;
}

extern "C" int cutdet_combine_adjacent(const cutdet_run_table *table, int64_t *n_runs, cutdet_stream_t stream) {
    if (int rc = check_table(table, "combine_adjacent")) return rc;
    CUTDET_REQUIRE(n_runs, "combine_adjacent: null n_runs");
    {
        KernelScope scope("combine_adjacent_kernel", as_stream(stream));
        combine_adjacent_kernel<<<1, 32, 0, as_stream(stream)>>>(*table, n_runs);
    }
    CUTDET_LAUNCH_CHECK("combine_adjacent_kernel");
    return CUTDET_OK;
}

extern "C" int cutdet_stitch_shards(const cutdet_run_table *src, int n_shards, int64_t shard_capacity, const int64_t *n_runs,
                                    const int64_t *offsets, const cutdet_run_table *dst, int64_t *n_out,
                                    cutdet_stream_t stream) {
    if (int rc = check_table(src, "stitch_shards(src)")) return rc;
    if (int rc = check_table(dst, "stitch_shards(dst)")) return rc;
    CUTDET_REQUIRE(n_shards >= 1 && shard_capacity > 0 && n_runs && offsets && n_out, "stitch_shards: bad argument");
    CUTDET_REQUIRE(src->capacity >= (int64_t)n_shards * shard_capacity, "stitch_shards: src smaller than n_shards * shard_capacity");
    {
        KernelScope scope("stitch_kernel", as_stream(stream));
        stitch_kernel<<<1, 256, 0, as_stream(stream)>>>(*src, n_shards, shard_capacity, n_runs, offsets, *dst, n_out);
    }
    CUTDET_LAUNCH_CHECK("stitch_kernel");
    return CUTDET_OK;
}

extern "C" size_t cutdet_shard_pack_bytes(int64_t capacity) { return capacity > 0 ? PACK_HEADER + (size_t)capacity * sizeof(PackedRow) : 0; }

extern "C" int cutdet_shard_pack(const cutdet_run_table *table, const int64_t *n_runs, int64_t n_frames, int64_t capacity,
                                 void *packed, cutdet_stream_t stream) {
    if (int rc = check_table(table, "shard_pack")) return rc;
    CUTDET_REQUIRE(n_runs && packed && capacity > 0 && n_frames >= 0, "shard_pack: bad argument");
    CUTDET_REQUIRE(reinterpret_cast<uintptr_t>(packed) % 8 == 0, "shard_pack: the packed buffer must be 8-byte aligned");
    const long long rows = capacity < table->capacity ? capacity : table->capacity;
    {
        KernelScope scope("shard_pack_kernel", as_stream(stream));
        shard_pack_kernel<<<(unsigned)std::min<long long>(ceil_div(rows, 256), 64), 256, 0, as_stream(stream)>>>(
            *table, n_runs, (long long)n_frames, (long long)rows, reinterpret_cast<uint8_t *>(packed));
    }
    CUTDET_LAUNCH_CHECK("shard_pack_kernel");
    return CUTDET_OK;
}

extern "C" int cutdet_stitch_packed(const void *gathered, int n_shards, int64_t capacity, const cutdet_run_table *dst, int64_t *n_out,
                                    int64_t *total_frames, cutdet_stream_t stream) {
    if (int rc = check_table(dst, "stitch_packed(dst)")) return rc;
    CUTDET_REQUIRE(gathered && n_shards >= 1 && capacity > 0 && n_out, "stitch_packed: bad argument");
    CUTDET_REQUIRE(reinterpret_cast<uintptr_t>(gathered) % 8 == 0, "stitch_packed: the gathered buffer must be 8-byte aligned");
    {
        KernelScope scope("stitch_packed_kernel", as_stream(stream));
        stitch_packed_kernel<<<1, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint8_t *>(gathered), n_shards, (long long)capacity,
                                                               *dst, n_out, total_frames);
    }
    CUTDET_LAUNCH_CHECK("stitch_packed_kernel");
    return CUTDET_OK;
}
