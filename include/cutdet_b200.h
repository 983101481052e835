/*
 * cutdet_b200.h -- C ABI of libcutdet_b200.so, the B200 (sm_100a) implementation of the
 * per-frame hot path of play4honor/Cut-Detection:
 *
 *     decoded frames -> resize/normalise -> CNN -> (max logit, label) -> run table -> smoothed segments
 *
 * The reference has no FFI of its own (it is pure Python over PyTorch/OpenCV); every entry point
 * below names the reference Python interface it stands in for (file:line under /root/reference).
 * The Python side of this repo (cut-detection_b200/frameID, a mirror of the reference's `frameID`
 * package) binds these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a CUTDET_E* code; cutdet_last_error() gives the text
 *     of the calling thread's last failure.  Nothing is ever computed on the CPU as a fallback.
 *   - pointers named *_dev are device pointers on the current CUDA device, *_host are host pointers.
 *     The caller owns every buffer it passes.  `stream` is a cudaStream_t (NULL = default stream);
 *     calls are asynchronous with respect to the host unless stated otherwise.
 *   - images are row-major; `frames` are uint8 BGR HWC exactly as cv2.VideoCapture.read() returns them.
 *   - class ids follow frameID/data.py:116  {a22: 0, ez: 1, b: 2}.
 */
#ifndef CUTDET_B200_H
#define CUTDET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUTDET_ABI_VERSION 2

#if defined(__GNUC__)
#define CUTDET_API __attribute__((visibility("default")))
#else
#define CUTDET_API
#endif

enum {
    CUTDET_OK = 0,
    CUTDET_EINVAL = 1,        /* bad argument (shape, null pointer, ...)                        */
    CUTDET_ECUDA = 2,         /* a CUDA runtime/driver call failed; see cutdet_last_error()     */
    CUTDET_EUNSUPPORTED = 3,  /* configuration outside what the kernels implement               */
    CUTDET_ECAPACITY = 4,     /* caller-provided table/workspace too small                      */
    CUTDET_ELONE_ORPHAN = 5   /* glue_orphans on a single orphan run: the reference raises
                                 IndexError here (frameID/segmentation.py:110-113)               */
};

typedef void *cutdet_stream_t;

CUTDET_API int cutdet_abi_version(void);
CUTDET_API const char *cutdet_last_error(void);
/* Fails with CUTDET_EUNSUPPORTED unless the current device is compute capability 10.x. */
CUTDET_API int cutdet_device_check(int *sm_count, int *cc_major, int *cc_minor);

/* Measurement hooks (bench.py): number of kernels this library has launched since it was loaded; and a per-kernel
 * CUDA-event profiler -- between begin and end every launch is bracketed by events on its own stream, end() drains
 * the device and writes {"kernel_name": {"launches": n, "ms": total}, ...} as JSON.                               */
CUTDET_API long long cutdet_launch_count(void);
CUTDET_API int cutdet_profile_begin(void);
CUTDET_API int cutdet_profile_end(char *json_out, size_t capacity);

/* ------------------------------------------------------------------------------------------------
 * K1  frame preprocessing            replaces VideoDataset.__init__/__next__, frameID/data.py:197-228
 * ------------------------------------------------------------------------------------------------ */

/* (new_width, new_height) = (resize, int(height * (resize / width)))       frameID/data.py:199-202 */
CUTDET_API int cutdet_target_size(int width, int height, int resize, int *new_width, int *new_height);

/* Geometry of one resize: the fixed-point tap tables cv2.resize(INTER_LINEAR) uses for uint8 images,
 * built on the host and kept on the device.  Created once per (h, w, H2, W2); thread-safe to share. */
typedef struct cutdet_resize_plan cutdet_resize_plan;
CUTDET_API int cutdet_resize_plan_create(int src_h, int src_w, int dst_h, int dst_w, cutdet_resize_plan **plan);
CUTDET_API void cutdet_resize_plan_destroy(cutdet_resize_plan *plan);
/* Source rows the resize actually reads (sorted, unique).  n_rows_out receives the count; rows_host may
 * be NULL to query it.  A host-side caller can copy just these rows to the device (row-compacted frames). */
CUTDET_API int cutdet_resize_plan_rows(const cutdet_resize_plan *plan, int *rows_host, int *n_rows_out);
/* The same list from the geometry alone: pure host arithmetic, no CUDA call (a frame source can start gathering rows --
 * e.g. fork its decoder processes -- before the process has a CUDA context).  rows_host may be NULL to query the count
 * (at most src_h).                                                                                                   */
CUTDET_API int cutdet_resize_rows(int src_h, int src_w, int dst_h, int dst_w, int *rows_host, int *n_rows_out);

/* Layout of a batch of source frames in device memory.  `row_map_compact` != 0 says the buffer holds only
 * the rows listed by cutdet_resize_plan_rows(), in that order (row r of the list at row_pitch * r).      */
typedef struct {
    const uint8_t *frames_dev; /* [B] frames, each `frame_stride` bytes apart                             */
    int64_t frame_stride;      /* bytes between consecutive frames                                        */
    int64_t row_pitch;         /* bytes between consecutive stored rows (>= 3 * src_w)                    */
    int batch;
    int row_map_compact;
} cutdet_frames;

/* Host -> device copy of ONLY the source rows the resize reads, as a handful of strided 2-D DMA copies (one at 720p,
 * where the rows are 5y+2; four at 1080p).  frames_host: [batch] frames of src_h rows, `row_pitch` bytes per row,
 * `frame_stride` bytes apart (pinned memory for an asynchronous copy).  dst_dev receives row-compacted frames
 * [batch][n_rows][3*src_w] -- pass them on with cutdet_frames.row_map_compact = 1.  Asynchronous on `stream`.      */
CUTDET_API int cutdet_upload_frames(const cutdet_resize_plan *plan, const uint8_t *frames_host, int batch,
                                    int64_t frame_stride, int64_t row_pitch, uint8_t *dst_dev,
                                    cutdet_stream_t stream, int64_t *bytes_copied);

/* uint8 BGR HWC -> float32 RGB CHW in [0,1]: bit-exactly what VideoDataset yields (data.py:220-228),
 * stacked to [B,3,H2,W2] as default_collate does (segment_video.py:29).                                  */
CUTDET_API int cutdet_preprocess_f32(const cutdet_resize_plan *plan, const cutdet_frames *src, float *out_nchw_dev,
                          cutdet_stream_t stream);
/* Measurement aid (tools/time_k1.py): which K1 kernel the two entry points below launch.  0 = the library chooses (default),
 * 1 = the one-thread-per-pixel kernel (any alignment), 2 = the row kernel (128-bit staged rows; 16-byte aligned frames),
 * 3 = the quad kernel (two-tap resizes, four adjacent output pixels per thread; falls back where it does not apply).
 * Process-wide; the results are bit-identical either way.                                                          */
CUTDET_API int cutdet_debug_k1_kernel(int mode);
/* uint8 BGR HWC -> resized uint8 BGR HWC [B,H2,W2,3]: bit-exactly cv2.resize(..., INTER_LINEAR).          */
CUTDET_API int cutdet_preprocess_u8(const cutdet_resize_plan *plan, const cutdet_frames *src, uint8_t *out_hwc_dev,
                         cutdet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K2/K3  the classifier             replaces FrameConvNet + FrameLinearNet, frameID/net.py:71-189,
 *                                   and load_and_glue_nets, frameID/net.py:193-217
 * ------------------------------------------------------------------------------------------------ */
typedef struct cutdet_net cutdet_net;

typedef struct {
    int input_channels;    /* FrameConvNet(input_channels=3, ...)            net.py:77        */
    int hidden_channels;   /* conv_channels                                  net.py:78        */
    int n_conv_layers;     /*                                                net.py:78        */
    int avg_pool_size;     /* AdaptiveAvgPool2d(average_pool_size)           net.py:79,88     */
    int n_fc_layers;       /* FrameLinearNet(n_layers, ...)                  net.py:148-151   */
    int fc_input_size;     /* must equal hidden_channels * avg_pool_size^2                   */
    int fc_hidden_size;
    int fc_output_size;
} cutdet_net_config;

CUTDET_API int cutdet_net_create(const cutdet_net_config *cfg, cutdet_net **net);
CUTDET_API void cutdet_net_destroy(cutdet_net *net);
/* Parameters come as host float32 arrays in PyTorch state_dict layout (eval-mode BatchNorm:
 * running statistics).  conv weight [Cout,Cin,3,3]; fc weight [out,in]; bn_* are NULL for the last
 * FC layer, which has no BatchNorm (net.py:164-167).                                                 */
CUTDET_API int cutdet_net_set_conv_layer(cutdet_net *net, int layer, const float *weight_host, const float *bias_host,
                              const float *bn_weight_host, const float *bn_bias_host,
                              const float *bn_mean_host, const float *bn_var_host, float bn_eps);
CUTDET_API int cutdet_net_set_fc_layer(cutdet_net *net, int layer, const float *weight_host, const float *bias_host,
                            const float *bn_weight_host, const float *bn_bias_host,
                            const float *bn_mean_host, const float *bn_var_host, float bn_eps);
/* Packs the parameters into the kernels' layouts and uploads them.  Synchronous. */
CUTDET_API int cutdet_net_finalize(cutdet_net *net);
/* 1 if the tcgen05 tensor-core kernels cover this architecture and input size, else 0 (the generic
 * CUDA-core kernels run instead).                                                                   */
CUTDET_API int cutdet_net_uses_tensor_cores(const cutdet_net *net, int height, int width);

/* Switches of the tensor-core path, per net (the library reads nothing from the environment).  They change how the
 * work is scheduled or which accumulator precision layer 1 uses, never what is computed; the workspace size depends on
 * SUB_BATCH / GROUP_FRAMES, so set them before asking for it.                                                       */
enum {
    CUTDET_OPT_CONV1_ACC32 = 1,  /* fused frames kernel: fp32 accumulators in layer 1 (default 0: fp16 accumulators)    */
    CUTDET_OPT_SUB_BATCH = 2,    /* frames per conv1/conv2 pass (default 0 = one frame per SM of a B200: 148)           */
    CUTDET_OPT_GROUP_FRAMES = 3, /* frames per conv12_frames / conv3 launch (default 0 = 4144 = 28 per SM)                    */
    CUTDET_OPT_NO_PDL = 4,       /* 1: ordinary launches instead of programmatic dependent launch                       */
    CUTDET_OPT_CONV1_GRID = 5,   /* test hook: cap on the fused conv1 grid (several frames per CTA); 0 = no cap         */
    CUTDET_OPT_RING_CAP = 8,     /* conv12_frames, resizes that read two source rows per output row (bilinear, 2x2): 0 (default) = a
                                    768-position operand ring where the raw-row ring would otherwise hold fewer than 16 slots, the
                                    freed 24 KB go to the raw-row ring (nine 11.5 KB slots at 1080p instead of six); 1 = always the
                                    full 1,024-position operand ring; 2 = always the smaller one.  Same bits either way.       */
    CUTDET_OPT_SRC_PREFETCH = 9, /* experiment: conv12_frames' TMA loaders ask for a slot's source rows in the L2
                                    (cp.async.bulk.prefetch.tensor) before they wait for the slot: 0 (default) = never, 1 = for resizes
                                    with two source rows per output row, 2 = with every tensor map of the source rows.  A cache
                                    hint: same bits; measured neutral once the raw ring has nine slots (profiles/README.md).      */
    CUTDET_OPT_L2_PERSIST = 7,   /* experiment: 1 = conv12_frames marks its layer-1 slots as a persisting window of the L2 (sets the
                                    context's cudaLimitPersistingL2CacheSize to the device maximum on first use)               */
    CUTDET_OPT_CONV1_VARIANT = 6 /* which kernels run layers 1 and 2 of the fused frames path; all give the same bits.
                                    0 (default) and 2: ONE kernel takes a frame through K1 + layer 1 + layer 2 per CTA
                                    (conv12_frames_kernel; where it does not apply -- fp32 accumulators requested -- the
                                    two-kernel path runs); 3: the two-kernel path (conv1_fused_tc + conv2_tc per 148-frame
                                    sub-batch); 1: the two-kernel path with the experimental conv1 kernel of two alternating
                                    epilogue sets (profiles/README.md, round 2; kept for same-box A/B runs)                 */
};
CUTDET_API int cutdet_net_set_option(cutdet_net *net, int option, int value);
CUTDET_API int cutdet_net_get_option(const cutdet_net *net, int option, int *value);

/* Bytes of device scratch the forward pass needs for `batch` inputs of height x width. */
CUTDET_API int cutdet_net_workspace_bytes(const cutdet_net *net, int batch, int height, int width, size_t *bytes);

/* net(x): x float32 [B,Cin,H,W] (RGB in [0,1]) -> raw logits float32 [B,fc_output_size]
 * (segment_video.py:45; no softmax anywhere in the reference).                                        */
CUTDET_API int cutdet_net_forward_f32(cutdet_net *net, const float *x_nchw_dev, int batch, int height, int width,
                           float *logits_dev, void *workspace_dev, size_t workspace_bytes,
                           cutdet_stream_t stream);
/* The same forward pass with every BatchNorm in TRAINING mode -- normalised with the mean and biased variance of this batch, as
 * nn.BatchNorm2d/1d do on a module that was never put in .eval() -- which is how the reference's contrastive script runs its
 * encoder (training_scripts/learn_contrasts.py:100-107: no .eval() anywhere).  Forward only: running statistics are not
 * updated and nothing is recorded for a backward pass.  batch >= 2.  use_tensor_cores != 0: the tcgen05 kernels with the
 * identity affine, then per-layer statistics/affine kernels (architectures the tensor-core path covers, batches of up to 148
 * frames -- the statistics span the batch); otherwise, and for everything else, float32 CUDA-core kernels.            */
CUTDET_API int cutdet_net_forward_f32_batchstats(cutdet_net *net, const float *x_nchw_dev, int batch, int height, int width,
                           float *out_dev, void *workspace_dev, size_t workspace_bytes, int use_tensor_cores,
                           cutdet_stream_t stream);
/* Fused entry: decoded frames in, logits out (K1 feeds the conv stack directly). */
CUTDET_API int cutdet_net_forward_frames(cutdet_net *net, const cutdet_resize_plan *plan, const cutdet_frames *src,
                              float *logits_dev, void *workspace_dev, size_t workspace_bytes,
                              cutdet_stream_t stream);
/* ONE layer of a finalized net, float32 CUDA-core kernels (the per-layer modules of the reference are callable on
 * their own):
 *   CNNLayer.forward, frameID/net.py:33-40   x [B,Cin,H,W] -> conv3x3(p1) -> ReLU -> MaxPool(3) -> BatchNorm -> [B,Cout,H/3,W/3]
 *   FCLayer.forward,  frameID/net.py:62-68   x [B,in] -> Linear -> (ReLU if relu) -> BatchNorm1d -> [B,out]
 * bn_mode: 0 = no BatchNorm (nn.Identity), 1 = running statistics (.eval()), 2 = statistics of this batch (training
 * mode, forward only).  A lone FCLayer with a BatchNorm is a one-layer FC-only net whose layer was given BN parameters. */
CUTDET_API int cutdet_net_forward_conv_layer(cutdet_net *net, int layer, const float *x_nchw_dev, int batch, int height,
                                  int width, float *out_nchw_dev, int bn_mode, cutdet_stream_t stream);
CUTDET_API int cutdet_net_forward_fc_layer(cutdet_net *net, int layer, const float *x_dev, int batch, float *out_dev,
                                int relu, int bn_mode, cutdet_stream_t stream);

/* Debug aid (tools/timeline.py, bench.py's roofline.phases): while armed (kernel 1 = conv1_fused_tc, 2 = conv2_tc, 3 =
 * conv12_frames; 0 or a null buffer disarms), CTA 0 of every launch of that kernel over at least 148 frames writes clock stamps
 * into the caller's device buffer of >= 4096 int64 entries (kernels 1 and 3: entries 2048 + 2 b, 2049 + 2 b = %globaltimer at
 * entry / exit of CTA b; kernel 3: [0] = clock64 at its start, [1 + 3 i], [2 + 3 i], [3 + 3 i] = layer 1 set up / layer 1 done /
 * layer 2 done of CTA 0's i-th frame, and per-role stamps of its third frame, see csrc/conv_tc.cu).
 * The library allocates, copies and synchronises nothing for it.                                                          */
CUTDET_API int cutdet_net_debug_timeline(cutdet_net *net, int kernel, long long *stamps_dev, size_t n_entries);

/* Intermediate activations of the last forward, converted to float32 NCHW (test hook):
 * layer in [0, n_conv_layers) is the output of that CNNLayer.                                         */
CUTDET_API int cutdet_net_debug_conv_output(cutdet_net *net, int layer, int batch, int height, int width,
                                 const void *workspace_dev, float *out_nchw_dev, cutdet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Contrastive objective (forward)   replaces ContrastiveLoss.forward, frameID/metrics.py:23-47
 * ------------------------------------------------------------------------------------------------ */
/* x [2*pairs, dim] float32 (first half = view 1, second half = view 2 of the same images, learn_contrasts.py:104) ->
 * *loss_dev (mean over pairs of the two cross entropies) and, if not null, logits_ab_dev [pairs, pairs].          */
CUTDET_API size_t cutdet_contrastive_loss_workspace_bytes(int pairs);
CUTDET_API int cutdet_contrastive_loss(const float *x_dev, int pairs, int dim, float temperature, int h_norm, float *loss_dev,
                           float *logits_ab_dev, void *workspace_dev, size_t workspace_bytes, cutdet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Supervised objective (forward)    replaces torch.nn.CrossEntropyLoss(reduction="sum") and the per-class accuracy counters
 *                                   of the validation loop, training_scripts/supervised_training.py:132, 148, 186-193
 * ------------------------------------------------------------------------------------------------ */
/* logits [n, n_classes] float32, labels [n] int64 -> *loss_dev = sum_i -log_softmax(logits_i)[label_i] and, if not null,
 * correct_dev[c] = #{i: label_i = c and argmax(logits_i) = c}, total_dev[c] = #{i: label_i = c} (int64 [n_classes], first-index
 * argmax as torch.max).  workspace_dev: cutdet_cross_entropy_workspace_bytes(n_classes) bytes, 8-byte aligned.  bad_label_host:
 * NULL = fully asynchronous (out-of-range labels are skipped); else the call synchronises the stream and returns CUTDET_EINVAL if
 * a label was outside [0, n_classes) (torch raises "Target out of bounds").  Forward only.                                      */
CUTDET_API size_t cutdet_cross_entropy_workspace_bytes(int n_classes);
CUTDET_API int cutdet_cross_entropy_sum(const float *logits_dev, const int64_t *labels_dev, int64_t n, int n_classes, float *loss_dev,
                             int64_t *correct_dev, int64_t *total_dev, void *workspace_dev, size_t workspace_bytes,
                             int *bad_label_host, cutdet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K4  per-frame decision            replaces torch.max(scores, dim=1), frameID/segmentation.py:37
 * ------------------------------------------------------------------------------------------------ */
/* scores [N,C] float32 -> top[N] (max logit), labels[N] (first index of the max, as torch does on CPU). */
CUTDET_API int cutdet_argmax(const float *scores_dev, int64_t n_frames, int n_classes, uint8_t *labels_dev,
                  float *top_dev, cutdet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K5  run-length encoding           replaces Segmentation.__init__, frameID/segmentation.py:39-60
 * ------------------------------------------------------------------------------------------------ */
/* A run table in device memory, struct-of-arrays, `capacity` rows.  Runs are appended in frame order.
 * start_frames[s] = end_frames[s-1] + 1, run_lengths[s] = end - start + 1,
 * score_means[s] = float32(sum of the run's max logits / run length).                                 */
typedef struct {
    int64_t *end_frames_dev;
    int64_t *start_frames_dev;
    int64_t *run_lengths_dev;
    int32_t *frame_types_dev;
    float *score_means_dev;
    double *score_sums_dev;   /* exact-ish (float64) sums, so runs cut by a shard edge can be re-joined */
    int64_t capacity;
} cutdet_run_table;

/* Streaming state (device memory, cutdet_rle_state_bytes() bytes, zero-initialised by cutdet_rle_reset):
 * the still-open last run is carried from one call to the next, so a long video can be encoded chunk
 * by chunk in frame order.                                                                            */
CUTDET_API size_t cutdet_rle_state_bytes(void);
CUTDET_API int cutdet_rle_reset(void *state_dev, cutdet_stream_t stream);
/* Append `n_frames` more frames (labels + max logits). */
CUTDET_API int cutdet_rle_append(void *state_dev, const uint8_t *labels_dev, const float *top_dev, int64_t n_frames,
                      const cutdet_run_table *table, cutdet_stream_t stream);
/* Close the open run; writes the number of runs to *n_runs_dev (device int64).  After this the table holds
 * exactly the five columns of Segmentation.te.                                                         */
CUTDET_API int cutdet_rle_finish(void *state_dev, const cutdet_run_table *table, int64_t *n_runs_dev,
                      cutdet_stream_t stream);
/* Synchronous helper: copies the run count (and the overflow flag) to the host after `stream` drains.
 * Returns CUTDET_ECAPACITY if the table overflowed.                                                    */
CUTDET_API int cutdet_rle_count(const void *state_dev, int64_t *n_runs_host, cutdet_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K6  segment smoothing             replaces Segmentation.glue_orphans / combine_adjacent_segments,
 *                                   frameID/segmentation.py:91-183 (incl. _find_orphans :12-17 and the
 *                                   mean update of _update_neighbor :69-89, evaluated as written)
 * ------------------------------------------------------------------------------------------------ */
/* In place on the first *n_runs_dev rows of `table` (start/end/length/type/mean columns; sums untouched);
 * *n_runs_dev is updated.  *status_dev (device int32) receives CUTDET_OK or CUTDET_ELONE_ORPHAN.
 * Exact ties between orphan means are broken towards the lowest run index (the reference's
 * torch.argsort is unstable, so its tie order is unspecified); a NaN mean sorts after every number,
 * as torch.argsort places it.
 * workspace_dev: cutdet_glue_orphans_workspace_bytes(table->capacity) bytes of device scratch, 16-byte
 * aligned, owned by the caller and private to this call until `stream` has run it (the library keeps
 * no state of its own: calls on different streams or devices need different workspaces).              */
CUTDET_API size_t cutdet_glue_orphans_workspace_bytes(int64_t capacity);
CUTDET_API int cutdet_glue_orphans(const cutdet_run_table *table, int64_t *n_runs_dev, int real_threshold,
                        int blank_threshold, int32_t *status_dev, void *workspace_dev, size_t workspace_bytes,
                        cutdet_stream_t stream);
CUTDET_API int cutdet_combine_adjacent(const cutdet_run_table *table, int64_t *n_runs_dev, cutdet_stream_t stream);

/* Joins the run tables of consecutive shards (time ranges) into one: `src` holds `n_shards` tables
 * back to back, shard i occupying rows [i * shard_capacity, i * shard_capacity + n_runs[i]) with frame
 * numbers LOCAL to the shard; frame_offsets_dev[i] is the shard's first global frame.  Runs that touch
 * across a shard edge with the same type are merged (sums and lengths add).  Output: global frame numbers;
 * *n_runs_out_dev may exceed dst->capacity, in which case the surplus rows were dropped. */
CUTDET_API int cutdet_stitch_shards(const cutdet_run_table *src, int n_shards, int64_t shard_capacity,
                         const int64_t *n_runs_dev, const int64_t *frame_offsets_dev,
                         const cutdet_run_table *dst, int64_t *n_runs_out_dev, cutdet_stream_t stream);

/* The exchange step of the multi-GPU path in two launches around ONE all-gather (the reference is
 * single-process).  cutdet_shard_pack writes this shard's table into a fixed-size packed buffer of
 * cutdet_shard_pack_bytes(capacity) bytes: {int64 n_runs, int64 n_frames} then `capacity` rows of 40 bytes {int64 end, start,
 * length; float64 sum; int32 type; float32 mean} with LOCAL frame numbers.  After the all-gather (ncclAllGather of equal-size
 * buffers, rank order = time order) cutdet_stitch_packed reads the gathered buffer as it is -- run counts and frame offsets
 * come from the headers, on the device -- and writes the joined table with global frame numbers.  No host synchronisation on
 * either side.  *n_runs_out_dev: the joined run count; > dst->capacity if rows were dropped; -(1 + k) if some shard had
 * k > capacity runs (nothing useful was written: gather again with a larger capacity).                                      */
CUTDET_API size_t cutdet_shard_pack_bytes(int64_t capacity);
CUTDET_API int cutdet_shard_pack(const cutdet_run_table *table, const int64_t *n_runs_dev, int64_t n_frames, int64_t capacity,
                      void *packed_dev, cutdet_stream_t stream);
CUTDET_API int cutdet_stitch_packed(const void *gathered_dev, int n_shards, int64_t capacity, const cutdet_run_table *dst,
                         int64_t *n_runs_out_dev, int64_t *total_frames_out_dev, cutdet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CUTDET_B200_H */
