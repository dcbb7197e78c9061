/*
 * stemseg_b200 -- C ABI of the B200-native (sm_100a) STEm-Seg hot path.
 *
 * The reference (sabarim/STEm-Seg) is pure Python/PyTorch and has no FFI boundary of its own;
 * its seams for this path are Python call sites.  Every entry point below names the reference call site whose
 * arithmetic it replaces (paths relative to the reference root).  The Python host layer in stemseg_b200/*.py
 * mirrors the reference's classes (same names, arguments, state_dict keys, error behaviour) on top of this ABI;
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every *device* pointer is documented as such; no allocation inside the
 *     library -- the caller owns every buffer, workspace sizes come from the *_workspace_bytes queries;
 *   - all work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*); functions return
 *     without synchronising unless stated;
 *   - return 0 on success, a negative STEMSEG_ERR_* code otherwise; stemseg_last_error() gives the message
 *     (thread-local);
 *   - tensors: "NDHWC" means [N][T][H][W][C] row-major with C innermost ("channels-last-3d").
 */
#ifndef STEMSEG_B200_H_
#define STEMSEG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STEMSEG_OK 0
#define STEMSEG_ERR_INVALID_ARGUMENT (-1)
#define STEMSEG_ERR_CUDA (-2)
#define STEMSEG_ERR_UNSUPPORTED (-3)
#define STEMSEG_ERR_WORKSPACE (-4)

#define STEMSEG_MAX_EMBEDDING_DIMS 16
#define STEMSEG_MAX_INSTANCES 64
#define STEMSEG_MAX_LOSS_INSTANCES 32

/* Message of the last failing call on this thread ("" if none). */
const char* stemseg_last_error(void);
/* ABI version of the loaded library (bumped on any signature change). */
int32_t stemseg_abi_version(void);
/* 0 if the current CUDA device is sm_100 (B200), else STEMSEG_ERR_UNSUPPORTED. Never falls back to anything. */
int32_t stemseg_check_device(void);

/* ------------------------------------------------------------------------------------------------------------
 * Sequential Gaussian-bandwidth clustering
 *   replaces SequentialClustering._process            stemseg/inference/clusterers.py:60-166
 *            ._get_next_instance_center               stemseg/inference/clusterers.py:168-175
 *            compute_distance / distances_to_prob     stemseg/inference/clusterers.py:53-58
 * One persistent cooperative kernel runs every iteration (masked grid-wide argmax of seediness -> centre ->
 * distances -> threshold -> label write) plus the secondary assignment on the device with no host round trip.
 *
 * Probability thresholds arrive in the distance domain: d_primary / d_secondary are the largest fp32 distances
 * d with exp(-0.5 d) > p  (clusterers.py:53-54,136-138,156-157; the map is monotone), computed on the host by
 * stemseg_prob_threshold_to_distance(); the kernel tests `d <= d_*` and contains no exp.
 *
 * Result layout of `meta` (device, 4-byte words, fetched by the caller with one D2H copy):
 *   word 0            K  = number of clusters created                    (len(instance_labels), clusterers.py:121)
 *   word 1            exit reason: 0 loop ran max_instances times, 1 no unassigned point left (clusterers.py:109),
 *                                  2 best seediness < min_seediness_prob (clusterers.py:116)
 *   word 2            number of points clustered (differs from n_points only with n_points_dev)
 *   word 3            reserved
 *   then int32  seed_index[max_instances]       index of the seed point of cluster k
 *   then float  centers[max_instances][E]       instance_centers (clusterers.py:124)
 *   then float  bandwidths[max_instances][E]    cat(bandwidth[seed], free_dim_bandwidths) (clusterers.py:119);
 *                                               instance_stds = sqrt(clamp(1/bw, 1e-8)) is left to the host
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct StemsegClusterParams {
    int64_t n_points;            /* N                                                                          */
    int32_t embedding_dims;      /* E  (1..STEMSEG_MAX_EMBEDDING_DIMS)                                          */
    int32_t n_free_dims;         /* f; bandwidths has E-f columns, the last f come from free_dim_bandwidths     */
    float free_dim_bandwidths[STEMSEG_MAX_EMBEDDING_DIMS]; /* 1/std^2, clusterers.py:100-102                   */
    float d_primary;             /* see above                                                                  */
    float d_secondary;
    float min_seediness_prob;    /* compared in fp32 like torch does (clusterers.py:116)                       */
    int32_t max_instances;       /* clusterers.py:36 (default 20), <= STEMSEG_MAX_INSTANCES                    */
    int64_t cluster_label_start; /* clusterers.py:60                                                           */
} StemsegClusterParams;

size_t stemseg_seq_cluster_meta_words(int32_t embedding_dims, int32_t max_instances);
int32_t stemseg_seq_cluster_workspace_bytes(const StemsegClusterParams* params, size_t* bytes);
/*
 * embeddings  device float [N][E]      (row-major, contiguous)       clusterers.py:61
 * bandwidths  device float [N][E-f]    already activated (exp*10, inference_model.py:148)
 * seediness   device float [N]         (the reference's [N,1] squeezed, clusterers.py:84)
 * labels      device int64 [N]  out    final labels, -1 = unassigned (clusterers.py:96,143,159)
 * primary     device int32 [N]  out    ordinal k of the cluster that claimed the point in the primary pass, -1 if
 *                                      none (instance_masks[k] == (primary == k), clusterers.py:145-146)
 * meta        device words      out    see above
 * n_points_dev optional device int32: the actual point count (<= params->n_points, which is then the capacity
 *             of the buffers); lets compaction -> gather -> clustering run back to back without a host round trip.
 *             meta word 2 reports the count that was clustered (0 points -> K = 0, exit reason 1).
 * Precondition params->n_points >= 1 (the N == 0 early return of clusterers.py:62-69 is host logic).
 */
int32_t stemseg_seq_cluster(const float* embeddings, const float* bandwidths, const float* seediness,
                            const StemsegClusterParams* params, const int32_t* n_points_dev, int64_t* labels,
                            int32_t* primary, void* meta, void* workspace, size_t workspace_bytes, void* stream);

/* Host helper: largest fp32 d >= 0 with fl32(exp(-0.5 d)) > fl32(p); -1 if none, +inf if every d qualifies. */
float stemseg_prob_threshold_to_distance(double prob_threshold);

/* ------------------------------------------------------------------------------------------------------------
 * Foreground compaction + gather
 *   replaces masks_to_coord_list                       stemseg/inference/online_chainer.py:11-22
 *            the gather in cluster_subsequence         stemseg/inference/online_chainer.py:258-281
 * stemseg_fg_compact: ordered stream compaction of a [T][H*W] uint8 mask (non-zero = foreground) into linear
 * voxel indices t*H*W + y*W + x, frame-major / row-major like torch.nonzero, plus per-frame counts.
 * stemseg_fg_gather: channel-first maps -> point-major rows for the n compacted voxels.
 * ---------------------------------------------------------------------------------------------------------- */
size_t stemseg_fg_compact_workspace_bytes(int64_t n_frames, int64_t frame_voxels);
/*
 * mask          device uint8 [T][HW]
 * indices       device int32 [T*HW] out (first total entries valid)
 * frame_counts  device int32 [T+1]  out (counts per frame, then the total)
 */
int32_t stemseg_fg_compact(const uint8_t* mask, int64_t n_frames, int64_t frame_voxels, int32_t* indices,
                           int32_t* frame_counts, void* workspace, size_t workspace_bytes, void* stream);
/* Same, with the foreground test `values > threshold` on an fp32 map [T][HW] fused in (seediness > 0.25,
 * stemseg/inference/main.py:93-103; foreground probability > 0.5, main.py:142-144). */
int32_t stemseg_fg_compact_threshold(const float* values, float threshold, int64_t n_frames, int64_t frame_voxels,
                                     int32_t* indices, int32_t* frame_counts, void* workspace,
                                     size_t workspace_bytes, void* stream);
/* Foreground from per-frame running sums: fg = upsample(sum / count[frame], (1, factor, factor)) > threshold on the
 * [T][factor*h][factor*w] grid -- the seediness / foreground-logit average over the sub-clips covering a frame
 * (stemseg/inference/main.py:93-103; inference_model.py:126-128,207), at full resolution when factor = 4
 * (--resize_embeddings: inference_model.py:55-61, online_chainer.py:128-140). */
int32_t stemseg_fg_compact_mean_threshold(const float* sum, const float* count, float threshold, int64_t n_frames,
                                          int32_t h, int32_t w, int32_t factor, int32_t* indices,
                                          int32_t* frame_counts, void* workspace, size_t workspace_bytes, void* stream);
/* dst[frame_ids[j]] += src[j] for the `frames` planes of one sub-clip; counts[frame_ids[j]] += 1 */
int32_t stemseg_frame_accumulate(float* dst, float* counts, const float* src, const int32_t* frame_ids, int32_t frames,
                                 int64_t plane, void* stream);
/* stemseg_fg_gather with the trilinear (1, factor, factor) up-sampling of online_chainer.py:128-140 fused in:
 * `indices` address the full-resolution grid [T][factor*h][factor*w], src is the [C][T][h][w] map; transform 1
 * activates (exp * 10) the corner values before interpolating, like the reference resizes activated bandwidths. */
int32_t stemseg_fg_gather_upsampled(const float* src, int64_t channel_stride, int32_t channels, int32_t h, int32_t w,
                                    int32_t factor, const int32_t* indices, int64_t n, const int32_t* n_dev,
                                    int32_t transform, float* dst, void* stream);
/*
 * src       device float [C][T*HW] channel-first map, channel stride `channel_stride` elements
 * indices   device int32 [n]
 * n_dev     optional device int32: actual number of points (<= n, n is then the capacity of indices / dst)
 * transform 0 = copy; 1 = exp(x) * 10, the bandwidth activation of inference_model.py:148 fused into the gather
 * dst       device float [n][C] out
 */
int32_t stemseg_fg_gather(const float* src, int64_t channel_stride, int32_t channels, const int32_t* indices,
                          int64_t n, const int32_t* n_dev, int32_t transform, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * 3-D decoder heads (SqueezingExpandDecoder x3)
 *   replaces the forward of  stemseg/modeling/embedding_decoder.py:101-145
 *                            stemseg/modeling/seediness_decoder.py:82-112
 *                            stemseg/modeling/semseg_decoder.py:91-116
 * Internal layout: activations are NDHWC bf16 "planes": plane 0 = bf16(x), plane 1 = bf16(x - plane0) when
 * planes == 2 (fp32-parity mode: three tensor-core products hi*hi + hi*lo + lo*hi), plane 0 only when
 * planes == 1 (bf16 mode).  Plane p of a tensor starts p * numel elements after plane 0.
 * ---------------------------------------------------------------------------------------------------------- */

/* NCTHW fp32 (any n/c/t strides, H*W contiguous) -> NDHWC bf16 planes.  Replaces the permute / stack copies of
 * restore_temporal_dimension (model_builder.py:84-99) and torch.stack(.., 2) (inference_model.py:112-119). */
int32_t stemseg_pack_activation(const float* src, int64_t stride_n, int64_t stride_c, int64_t stride_t, int32_t n,
                                int32_t c, int32_t t, int32_t hw, void* dst_planes, int32_t planes, void* stream);

/* nn.Conv3d / 1x1x1 weight [cout][cin_total][taps] fp32 (state_dict layout, taps = 27 or 1) -> K-major bf16 planes
 * [rows_total][taps][cin_count] at rows [row_begin, row_begin+cout), input channels [cin_begin, +cin_count).
 * Row offsets let several heads share one GEMM N dimension; channel ranges split the concat-conv weights
 * (embedding_decoder.py:68,74,80) into their upsampled / skip halves. */
int32_t stemseg_pack_conv_weight(const float* src, int32_t cout, int32_t cin_total, int32_t cin_begin,
                                 int32_t cin_count, int32_t taps, void* dst_planes, int32_t row_begin,
                                 int32_t rows_total, int32_t planes, void* stream);

/* Operand storage formats (the `planes` argument of every function that reads or writes packed operands):
 *   1  one bf16 plane (bf16 mode, one tensor-core product per MAC)
 *   2  two bf16 planes hi + lo (fp32-parity mode, three products per MAC)
 *  17  one fp16 plane (STEMSEG_PLANES_FP16): one product per MAC with 11-bit operands.  An opt-in for block_8x /
 *      block_16x of the fp32-parity plan (profiles/r02_precision_ablation.json: inside 1e-4 on the shipped 8-frame
 *      widths but without margin on every golden, so the host layer keeps it off by default; never admissible for the
 *      4x layer or the merges).  Accepted by stemseg_pack_activation, stemseg_to_planes, stemseg_pack_conv_weight,
 *      stemseg_norm_relu_pool and stemseg_conv3d_bf16_planes. */
#define STEMSEG_PLANES_FP16 17

typedef struct StemsegConvShape {
    int32_t n, t, h, w;      /* output volume == input volume (stride 1, zero padding 1 for kernel_size 3)        */
    int32_t cin, cout;       /* multiples of 32                                                                 */
    int32_t kernel_size;     /* 3 (3x3x3, embedding_decoder.py:21) or 1 (1x1x1 merge, embedding_decoder.py:68)  */
    int32_t planes;          /* 1 or 2                                                                          */
    int32_t split_k;         /* 1, or 3 / 9 / 27 (kernel_size 3): number of tap slices computed by separate CTAs;
                                out then holds split_k partial sums [split_k][n][t][h][w][cout] (bias in slice 0)
                                which the GroupNorm kernels add in a fixed order                                  */
    int32_t tiles_per_cta;   /* 0: persistent CTAs (one per SM) looping over the tiles; > 0: short-lived CTAs that own
                                this many consecutive tiles -- lets higher-priority kernels of concurrent branches
                                obtain SMs while a long layer runs                                                */
    int32_t out_bf16;        /* 1: `out` receives bf16 [n][t][h][w][cout] instead of fp32 (bf16 mode; split_k == 1 only):
                                halves the write of the conv output and the read of the GroupNorm-apply pass; the
                                GroupNorm statistics (stat_partial) are still taken from the fp32 accumulators        */
} StemsegConvShape;

/* Split that fills the SMs for a latency-bound (few-tile) layer on the current device; 1 for large layers. */
int32_t stemseg_conv3d_auto_split(const StemsegConvShape* shape);

/* out[n][t][h][w][cout] (fp32) = conv(act) + bias.  tcgen05 implicit GEMM, TMA im2col, fp32 accumulation in TMEM.
 * max_ctas > 0 caps the persistent grid (used to run independent branches concurrently on separate streams). */
int32_t stemseg_conv3d_bf16_planes(const void* act_planes, const void* weight_planes, const float* bias, float* out,
                                   float* stat_partial, const StemsegConvShape* shape, int32_t max_ctas, void* stream);
/* stat_partial (optional, split_k == 1 only): [n][cout][tiles][2] per-tile (sum, sum of squares) of every output
 * channel, written by the conv epilogue from the accumulator registers; tiles = stemseg_conv3d_tiles_per_sample().
 * stemseg_group_norm_finalize turns it into the GroupNorm affine table without re-reading the conv output. */
int32_t stemseg_conv3d_tiles_per_sample(const StemsegConvShape* shape);

/* GroupNorm statistics of an NDHWC fp32 tensor -> per-channel affine table scale_shift[n][c][2] with
 *   scale = rstd * gamma, shift = beta - mean * rstd * gamma   (biased variance, eps inside the sqrt; nn.GroupNorm(32, C),
 * model_builder.py:34).  Deterministic (fixed reduction order, final combination in double).
 * x may be `slices` split-K partial sums [slices][n][spatial][row_stride] (see StemsegConvShape.split_k): they are
 * added in a fixed order and THE SUM IS WRITTEN BACK INTO SLICE 0, so later passes read one slice.  row_stride >= c
 * is the element distance between voxels, so x can be a channel slice of a wider (multi-head) conv output. */
size_t stemseg_group_norm_workspace_bytes(int32_t n, int64_t spatial, int32_t c);
int32_t stemseg_group_norm_stats(float* x, int32_t row_stride, int32_t slices, int32_t n, int64_t spatial, int32_t c,
                                 int32_t channels_per_group, float eps, const float* gamma, const float* beta,
                                 float* scale_shift, float* mean_rstd, void* workspace, size_t workspace_bytes,
                                 void* stream);
/* mean_rstd (optional, [n][c/channels_per_group][2]): the group statistics themselves, saved for the backward pass */

/* Second half of stemseg_group_norm_stats for partial sums produced elsewhere (the conv epilogue): `partial` points
 * at channel 0 of the c channels to normalise inside a [n][c_total][chunks][2] buffer, partial_sample_stride =
 * c_total*chunks*2 floats. */
int32_t stemseg_group_norm_finalize(const float* partial, int64_t partial_sample_stride, int32_t chunks, int32_t n,
                                    int64_t spatial, int32_t c, int32_t channels_per_group, float eps,
                                    const float* gamma, const float* beta, float* scale_shift, float* mean_rstd,
                                    void* stream);

/* relu(x * scale + shift) [-> pooling] -> operand planes (embedding_decoder.py:22-24; common.py:8-24).
 * pool: 0 none, 1 AvgPool3d(3, stride=(2,1,1), padding=1) with divisor 27 (cfg POOL_TYPE "avg"), 2 MaxPool3d with the
 * same window (POOL_TYPE "max", model_builder.py:28-30).  scale_shift NULL = no normalisation (NormType Identity). */
int32_t stemseg_norm_relu_pool(const float* x, int32_t row_stride, int32_t slices, const float* scale_shift, int32_t n,
                               int32_t t, int32_t h, int32_t w, int32_t c, int32_t pool, void* dst_planes,
                               int32_t planes, void* stream);
/* Same pass over a bf16 conv output (StemsegConvShape.out_bf16); slices must be 1. */
int32_t stemseg_norm_relu_pool_bf16in(const void* x, int32_t row_stride, int32_t slices, const float* scale_shift, int32_t n,
                               int32_t t, int32_t h, int32_t w, int32_t c, int32_t pool, void* dst_planes,
                               int32_t planes, void* stream);

/* dst = z + trilinear_upsample(y_low, (t_scale, 2, 2), align_corners=False) -> bf16 planes; z is [n][t][h][w][c],
 * y_low [n][t/t_scale][h/2][w/2][c] (common.py:69-78; conv1x1(cat(up(x), f)) == up(W_a x) + W_b f). */
int32_t stemseg_upsample_add(const float* z, const float* y_low, int32_t n, int32_t t, int32_t h, int32_t w,
                             int32_t c, int32_t t_scale, void* dst_planes, int32_t planes, void* stream);

/* Output heads on x = z + upsample(y_low): n_out 1x1x1 outputs with activation[j] (0 identity, 1 tanh(0.25 v),
 * 2 sigmoid) and coordinate[j] (0 none, 1 t, 2 y, 3 x) offsets; out is channels-first [n][n_out][t][h][w] fp32
 * (embedding_decoder.py:131-145, embedding_utils.py:29-120, seediness_decoder.py:112, semseg_decoder.py:116). */
int32_t stemseg_head_output(const float* z, const float* y_low, int32_t n, int32_t t, int32_t h, int32_t w,
                            int32_t c, int32_t t_scale, const float* out_weight, const float* out_bias,
                            const int32_t* activation, const int32_t* coordinate, int32_t n_out, float time_scale,
                            float* out, void* stream);

/* Fused variant of the last merge + output heads (what the plan normally runs): the 1x1x1 conv_4 GEMM
 * (embedding_decoder.py:80,128-129) keeps its accumulator tile on the SM and its epilogue applies the output convs,
 * activations and coordinate offsets directly -- the [n,t,h,w,c3] merge result is never written to HBM.
 *   out[n][j][t][h][w] = act_j( W_out[j] . (W_b f4')[voxel] + up(p_low)[voxel][j] + b_j ) + coord_j
 * with p_low = stemseg_head_lowres(W_a x_8) = the output convs applied at the low resolution (they commute with the
 * trilinear upsampling).  Same argument meaning as stemseg_head_output. */
int32_t stemseg_head_lowres(const float* y_low, int64_t voxels, int32_t c, const float* out_weight, int32_t n_out,
                            float* p_low, void* stream);
int32_t stemseg_conv1x1_head_output(const void* act_planes, const void* weight_planes, const StemsegConvShape* shape,
                                    const float* p_low, int32_t t_scale, const float* out_weight,
                                    const float* out_bias, const int32_t* activation, const int32_t* coordinate,
                                    int32_t n_out, float time_scale, float* out, int32_t max_ctas, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Sub-clip stitch on the device
 *   replaces the K1*K2 mask reductions of OnlineChainer.associate_clusters      online_chainer.py:315-328
 *            the per-association torch.where relabelling                        online_chainer.py:219-224
 *            the per-frame unique / count loops of get_track_mask_idxes         online_chainer.py:94-117
 * table[ia][ib] counts the points with bin(a) == ia and bin(b) == ib, where bin(v) = 0 for v < 0 (outliers) and
 * v - base + 1 otherwise; values outside the table increment *out_of_range (caller error).  Intersection, sizes and
 * unions of every label pair follow from this one table (host-side Hungarian: stemseg_b200.chaining.
 * stitch_subsequences_device; fully device-resident: stemseg_stitch_subclip below).
 * ---------------------------------------------------------------------------------------------------------- */
int32_t stemseg_label_pair_histogram(const int64_t* a, const int64_t* b, int64_t n, int64_t a_base, int64_t b_base,
                                     int32_t na, int32_t nb, int32_t* table, int32_t* out_of_range, void* stream);
/* labels[i] = lut[labels[i] - base] for labels[i] >= base (negative labels and labels outside the table unchanged) */
int32_t stemseg_relabel_lut(int64_t* labels, int64_t n, int64_t base, const int64_t* lut, int32_t nlut, void* stream);

/* The whole sequential stitch of one sub-clip, resident on the device (no host round trip per sub-clip):
 *   replaces OnlineChainer.process for sub-clip i (online_chainer.py:162-236): TrackContainer.add_labels / get_labels,
 *            associate_clusters incl. scipy.optimize.linear_sum_assignment (online_chainer.py:291-343, :330),
 *            the relabelling of the non-overlap frames (:213-224), the meta-info update (:227-229) and the statistics
 *            of TrackContainer.get_track_mask_idxes (:94-117).
 * labels        [capacity] LOCAL labels of the sub-clip (clustered with cluster_label_start = 1; -1 = outlier), frames
 *               concatenated; rewritten in place to the labels OnlineChainer.process returns in subseq_labels_list
 *               (overlap frames: local + next_track_label - 1; other frames: associated track ids).
 * frame_counts  DEVICE int32 [n_frames] points per frame;  k_dev  DEVICE int32: number of clusters of the sub-clip.
 * frames / overlap  HOST int32 [n_frames]: global frame number of each slot, and 1 if that frame is shared with the
 *               previous sub-clip (already in the container).  is_first: sub-clip 0 (no association).
 * container     frame_labels int64 [num_frames][frame_capacity], frame_count int32 [num_frames] (-1 = absent).
 * state         int32 [4]: next_track_label (init 1), highest id (init 0), error flags (init 0), sub-clips done.
 * track_counts  int64 [max_labels + 1] (index = id + 1: slot 0 counts the outliers), span_lo / span_hi int32
 *               [max_labels + 1] first / last frame of every id (init 10000 / -1 like online_chainer.py:97).
 * meta_labels_out  DEVICE int64 [max_instances]: instance_labels of the sub-clip after association (-1 padding).
 * Error flags (state[2]): 1 label outside range, 2 overlap frames differ in size, 4 label sets overlap, 8 too many
 * labels, 16 frame already present, 32 assignment failed.  The label lists are ordered like CPython's
 * list(set(...) - {-1}) and the assignment is scipy's algorithm, so ties resolve exactly as in the reference. */
size_t stemseg_stitch_workspace_bytes(int32_t max_instances, int32_t n_frames, int32_t max_labels);
int32_t stemseg_stitch_subclip(int64_t* labels, int64_t capacity, const int32_t* frame_counts, const int32_t* k_dev,
                               const int32_t* frames, const int32_t* overlap, int32_t n_frames, int32_t is_first,
                               int32_t max_instances, int64_t* frame_labels, int32_t* frame_count,
                               int64_t frame_capacity, int32_t num_frames, int32_t* state, int64_t* track_counts,
                               int32_t* span_lo, int32_t* span_hi, int32_t max_labels, int64_t* meta_labels_out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Instance-mask writeback
 *   replaces label scatter -> one-hot -> x4 bilinear -> crop -> bilinear resize -> "> 0.5" -> condensation
 *            stemseg/inference/output_utils/davis.py:76-112 (youtube_vis.py:117-161, kitti_mots.py:101-166)
 * stemseg_rank_map_scatter: rank_map[indices[i]] = lut[labels[i]] into a zeroed uint8 map of map_elems voxels
 *   (lut: track id -> rank + 1 among the instances to keep, 0 = dropped; ids outside the table / negative -> 0).
 * stemseg_mask_writeback: out[f][y][x] = rank of the unique instance whose two-stage interpolated one-hot mask exceeds
 *   0.5 at that pixel (0 if none): rank_map [frames][h][w] is up-sampled by `upscale` (bilinear, align_corners=False),
 *   cropped to crop_h x crop_w and resized to out_h x out_w (bilinear), in the reference's fp32 evaluation order.
 * ---------------------------------------------------------------------------------------------------------- */
int32_t stemseg_rank_map_scatter(const int32_t* indices, const int64_t* labels, int64_t n, const uint8_t* lut,
                                 int32_t nlut, uint8_t* rank_map, int64_t map_elems, void* stream);
int32_t stemseg_mask_writeback(const uint8_t* rank_map, int32_t frames, int32_t h, int32_t w, int32_t upscale,
                               int32_t crop_h, int32_t crop_w, int32_t out_h, int32_t out_w, uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backward pass of the decoder heads (the reference relies on torch autograd through
 * nn.Conv3d / nn.GroupNorm / nn.AvgPool3d / F.interpolate, stemseg/training/main.py:188-201).
 * GEMM-shaped parts reuse the tcgen05 convolution kernel:
 *   dgrad  = stemseg_conv3d_bf16_planes(dy planes, weights packed by stemseg_pack_conv_weight_dgrad);
 *   wgrad  = stemseg_conv3d_wgrad on zero-padded TRANSPOSED planes (stemseg_transpose_pad), 27 K-major GEMMs whose B
 *            operand is read at the constant offset of the tap, followed by stemseg_wgrad_reduce.
 * Batch size 1 per call (the reference trains with MAX_SAMPLES_PER_GPU = 1, defaults.yaml:20).
 * ---------------------------------------------------------------------------------------------------------- */
/* W'[ci][taps-1-tap][co] = W[co][cin_begin+ci][tap] as K-major bf16 planes (rows = ci, K = taps*cout) */
int32_t stemseg_pack_conv_weight_dgrad(const float* src, int32_t cout, int32_t cin_total, int32_t cin_begin,
                                       int32_t cin_count, int32_t taps, void* dst_planes, int32_t planes, void* stream);
/* output heads: given grad_out [n][n_out][t][h][w] -> dx [n][t][h][w][c] (x = z + up(y_low)), d_weight [n_out][c],
 * d_bias [n_out] */
size_t stemseg_head_backward_workspace_bytes(int32_t c);
int32_t stemseg_head_backward(const float* z, const float* y_low, int32_t n, int32_t t, int32_t h, int32_t w, int32_t c,
                              int32_t t_scale, const float* out_weight, const float* out_bias,
                              const int32_t* activation, int32_t n_out, const float* grad_out, float* dx,
                              float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes, void* stream);
/* Training-path output heads over the merged feature x = z + up(y_low) kept in fp32 (csrc/head_train.cu):
 * stemseg_upsample_add_f32 forms x in place of z once per step; stemseg_head_output_x / stemseg_head_backward_x are the
 * streaming forward / backward of the 1x1x1 output convs + activations + coordinate offsets
 * (embedding_decoder.py:90-96,131-145; seediness_decoder.py:80,112; semseg_decoder.py:86-87,116). */
int32_t stemseg_upsample_add_f32(float* z, const float* y_low, int32_t n, int32_t t, int32_t h, int32_t w, int32_t c,
                                 int32_t t_scale, void* stream);
int32_t stemseg_head_output_x(const float* x, int32_t n, int32_t t, int32_t h, int32_t w, int32_t c,
                              const float* out_weight, const float* out_bias, const int32_t* activation,
                              const int32_t* coordinate, int32_t n_out, float time_scale, float* out, void* stream);
size_t stemseg_head_backward_x_workspace_bytes(int32_t c);
int32_t stemseg_head_backward_x(const float* x, int32_t n, int32_t t, int32_t h, int32_t w, int32_t c,
                                const float* out_weight, const float* out_bias, const int32_t* activation, int32_t n_out,
                                const float* grad_out, float* dx, float* d_weight, float* d_bias, void* workspace,
                                size_t workspace_bytes, void* stream);
/* adjoint of the trilinear (t_scale, 2, 2) up-sampling: d_low [n][t/t_scale][h/2][w/2][c] from d_high [n][t][h][w][c] */
int32_t stemseg_upsample_transpose(const float* d_high, int32_t n, int32_t t, int32_t h, int32_t w, int32_t c,
                                   int32_t t_scale, float* d_low, void* stream);
/* AvgPool3d(3,(2,1,1),1) (if pool) + ReLU backward: d_norm [n][t][h][w][c] from d_out [n][t_out][h][w][c] */
int32_t stemseg_pool_relu_backward(const float* d_out, const float* y, const float* scale_shift, int32_t n, int32_t t,
                                   int32_t h, int32_t w, int32_t c, int32_t pool, float* d_norm, void* stream);
/* GroupNorm backward: d_norm_to_dy is overwritten with dy; dgamma_dbeta [n][c][2]; group_terms [n][groups][2] scratch */
size_t stemseg_group_norm_backward_workspace_bytes(int32_t n, int64_t spatial, int32_t c);
int32_t stemseg_group_norm_backward(float* d_norm_to_dy, const float* y, const float* mean_rstd, const float* gamma,
                                    int32_t n, int64_t spatial, int32_t c, int32_t channels_per_group,
                                    float* dgamma_dbeta, float* group_terms, void* workspace, size_t workspace_bytes,
                                    void* stream);
/* GroupNorm backward with the fused tail: dy is written only as bf16 planes [P][n][spatial][c] (what dgrad / wgrad
 * consume) and its per-channel sums (the conv bias gradient) come out of the same pass; d_norm is left untouched */
size_t stemseg_group_norm_backward_planes_workspace_bytes(int32_t n, int64_t spatial, int32_t c);
int32_t stemseg_group_norm_backward_planes(const float* d_norm, const float* y, const float* mean_rstd,
                                           const float* gamma, int32_t n, int64_t spatial, int32_t c,
                                           int32_t channels_per_group, float* dgamma_dbeta, float* group_terms,
                                           void* dy_planes, int32_t planes, float* d_bias, void* workspace,
                                           size_t workspace_bytes, void* stream);
/* out[c] = sum over rows of x[rows][c] (bias gradients), deterministic */
size_t stemseg_channel_sum_workspace_bytes(int64_t rows, int32_t c);
int32_t stemseg_channel_sum(const float* x, int64_t rows, int32_t c, float* out, void* workspace, size_t workspace_bytes,
                            void* stream);
/* fp32 -> bf16 planes without activation (gradients entering a dgrad GEMM) */
int32_t stemseg_to_planes(const float* x, int64_t elems, void* dst_planes, int32_t planes, void* stream);
/* [t*h*w][c] (fp32, or bf16 planes when src_is_planes) -> zero-padded transposed bf16 planes [P][shifts][c][row_length]
 * with row_length = stemseg_transposed_row_length(t, h, w, pad) and element (t,y,x) at
 * ((t+pad)(h+2pad)+(y+pad))*pitch + x+pad, pitch = w (pad 0) or round_up(w+2, 8) (pad 1).  shifts = 3 (pad 1 only)
 * writes three copies shifted by dw = -1, 0, +1 along the row (copy[q] = x[q+dw]): a TMA load needs a 16-byte
 * aligned start coordinate in the contiguous dimension, so the dw part of a filter-tap offset is taken from the copy
 * and only the (dt, dh) part (a multiple of the pitch) from the coordinate. */
int64_t stemseg_transposed_row_length(int32_t t, int32_t h, int32_t w, int32_t pad);
int32_t stemseg_transpose_pad(const void* src, int32_t src_is_planes, int32_t t, int32_t h, int32_t w, int32_t c,
                              int32_t pad, int32_t shifts, void* dst_planes, int32_t planes, void* stream);
/* slices [k_splits][taps][cout][cin] fp32 partial weight gradients; see the section comment */
int32_t stemseg_wgrad_k_splits(int32_t cout, int32_t cin, int32_t t, int32_t h, int32_t w, int32_t kernel_size,
                               int32_t planes);
/* dyT_planes: transpose_pad with shifts 1; xT_planes: shifts 3 when kernel_size is 3, else 1 */
int32_t stemseg_conv3d_wgrad(const void* dyT_planes, const void* xT_planes, int32_t cout, int32_t cin, int32_t t,
                             int32_t h, int32_t w, int32_t kernel_size, int32_t planes, int32_t k_splits, float* slices,
                             void* stream);
/* Direct weight gradient (the default): both operands stay NDHWC bf16 planes [P][1][t][h][w][C] (dy from
 * stemseg_to_planes, x = the forward pass's input planes) and are consumed as MN-major tensor-core tiles; the tap
 * shift / zero padding is a TMA coordinate on the x operand.  slices [k_splits][taps][cout][cin]. */
int32_t stemseg_wgrad_direct_k_splits(int32_t cout, int32_t cin, int32_t t, int32_t h, int32_t w, int32_t kernel_size);
int32_t stemseg_conv3d_wgrad_direct(const void* dy_planes, const void* x_planes, int32_t cout, int32_t cin, int32_t t,
                                    int32_t h, int32_t w, int32_t kernel_size, int32_t planes, int32_t k_splits,
                                    float* slices, void* stream);
/* sum the slices into the state_dict layout dst[cout][cin_total][taps] at input-channel offset cin_begin */
int32_t stemseg_wgrad_reduce(const float* slices, int32_t n_slices, int32_t cout, int32_t ntaps, int32_t cin, float* dst,
                             int32_t cin_total, int32_t cin_begin, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Embedding loss of one training sequence: loss terms AND the gradient wrt the head output in one call
 *   replaces EmbeddingLoss.forward / compute_prob_map / compute_bandwidth_smoothness_loss
 *                                                     stemseg/modeling/losses/embedding_loss.py:35-185
 *            lovasz_hinge_flat / lovasz_grad          stemseg/modeling/losses/_lovasz.py:139-157, :18-31
 *            and torch autograd through them          stemseg/training/main.py:195-201
 * head_out [E+V][voxels] fp32 channels-first (embedding rows then variance rows, V = E - n_free_dims), seediness
 * [voxels], masks [n_instances][voxels] uint8 (0 / non-zero), ignore [voxels] uint8 or NULL -- all device pointers.
 * Per kept instance: masked mean centre and mean activated bandwidth (10 exp(var), free dims 1/std^2), Gaussian
 * probability of EVERY voxel, Lovasz hinge (full descending sort of the hinge errors, done by an on-device
 * bitonic sort of 64-bit keys), seediness regression to the detached probability, variance smoothness.
 * One sequence per call (the reference trains with MAX_SAMPLES_PER_GPU = 1, defaults.yaml:20); quirks preserved:
 * empty instances are dropped but kept slot n is paired with masks[n] (embedding_loss.py:83-87,128); no background
 * voxel at all -> NaN seediness term (mean of an empty tensor).
 * Outputs (device): losses[4] = {total (weighted, x w), lovasz, variance_smoothness, seediness};
 * d_head_out [E+V][voxels] and d_seediness [voxels] = gradient of losses[0] (every element written).
 * No host synchronisation; reductions accumulate in double (order-independent to ~1e-16).
 * ---------------------------------------------------------------------------------------------------------- */
size_t stemseg_embedding_loss_workspace_bytes(int64_t voxels, int32_t n_instances);
int32_t stemseg_embedding_loss(const float* head_out, const float* seediness, const uint8_t* masks,
                               const uint8_t* ignore, int64_t voxels, int32_t n_instances, int32_t embedding_dims,
                               int32_t n_free_dims, const float* free_dim_stds /* host, n_free_dims */,
                               float w_lovasz, float w_variance_smoothness, float w_seediness, float w, float* losses,
                               float* d_head_out, float* d_seediness, void* workspace, size_t workspace_bytes,
                               void* stream);
/* ------------------------------------------------------------------------------------------------------------
 * Losses of the semantic-segmentation head, loss AND gradient, one sequence per call
 *   replaces CrossEntropyLoss.forward          stemseg/modeling/losses/cross_entropy.py:13-49
 *            TrainingModel.compute_fg_loss     stemseg/modeling/model_builder.py:210-244
 * class_logits: channel c of voxel v at class_logits[c * class_stride + v] (the [T,cls,H,W] view the reference passes
 * has exactly this form), n_classes channels; fg_logits [voxels]; class_ids [voxels] int64 (must lie in
 * [0, n_classes), the reference raises otherwise); ignore [voxels] uint8 or NULL.  Either term is skipped when its
 * logits pointer is NULL.  losses[2] = {class loss (unweighted), foreground loss}; the gradients are those of
 * w_semseg * losses[0] + w_foreground * losses[1].  Quirk kept: the class loss is a plain mean over ALL voxels times
 * S / S (S = non-ignored voxels): ignored voxels count, S = 0 gives NaN.
 * ---------------------------------------------------------------------------------------------------------- */
size_t stemseg_semseg_loss_workspace_bytes(void);
int32_t stemseg_semseg_loss(const float* class_logits, int64_t class_stride, int32_t n_classes, const float* fg_logits,
                            const int64_t* class_ids, const uint8_t* ignore, int64_t voxels, float w_semseg,
                            float w_foreground, float* losses, float* d_class, int64_t d_class_stride, float* d_fg,
                            void* workspace, size_t workspace_bytes, void* stream);
/* x[i] *= *scalar (scalar in device memory: chain-rule factor of loss.backward() without a host round trip) */
int32_t stemseg_scale_by_device_scalar(float* x, int64_t n, const float* scalar, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused SGD step over a flat fp32 parameter buffer
 *   replaces torch.optim.SGD(..., momentum, weight_decay, nesterov).step()   stemseg/training/utils.py:199-202,
 *                                                                            stemseg/training/main.py:209
 * g' = grad * grad_scale + weight_decay * p;  buf = momentum * buf + g';  p -= lr * (nesterov ? g' + momentum * buf : buf)
 * grad_scale = 1 / world_size folds the data-parallel mean into the pass.  momentum_buf must start zeroed.
 * ---------------------------------------------------------------------------------------------------------- */
int32_t stemseg_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                         float weight_decay, float grad_scale, int32_t nesterov, void* stream);
/* Same step with {lr, momentum, weight_decay, grad_scale} read from DEVICE memory (float[4]) at run time: a CUDA graph
 * that captured this launch follows the learning-rate schedule (training/main.py:209-210 steps the scheduler every
 * iteration) by updating the array, without re-capture. */
int32_t stemseg_sgd_step_dev(float* param, const float* grad, float* momentum_buf, int64_t n, const float* hyper,
                             int32_t nesterov, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STEMSEG_B200_H_ */
