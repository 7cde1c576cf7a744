/*
 * hvpr_b200 — C ABI of the B200-native hybrid voxel-point encoding front end.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.
 * Every entry point
 *   - takes DEVICE pointers unless a parameter says "host";
 *   - enqueues all its work on `stream` (a cudaStream_t passed as void*), never synchronises the device,
 *     never allocates device memory and keeps NO mutable state: launch-shape choices travel with the call
 *     (HvprLaunchCfg), the only cached datum is the immutable SM count of each device -> safe inside CUDA-graph
 *     capture and re-entrant across streams, threads and devices;
 *   - returns HVPR_OK (0) or a negative HvprStatus; it never throws.
 * The caller (hvpr_b200/*.py through ctypes, or any C/C++ host) owns every buffer.
 *
 * Reference interfaces replaced (file:line under the reference tree):
 *   hvpr_voxelize        spconv VoxelGenerator.generate() as called at pcdet/datasets/processor/data_processor.py:50-67
 *                        + the collate of pcdet/datasets/dataset.py:159-166 (rows frame-major, coords [b,z,y,x])
 *   hvpr_pfn             PillarVFE_Scale.forward  pcdet/models/backbones_3d/vfe/pillar_vfe.py:184-221
 *                        (PillarVFE.forward :94-124 when scale_out == NULL), PFNLayer.forward :29-49
 *   hvpr_mem_attn        MemoryUnit_Agg.forward eval branch  pcdet/models/backbones_2d/map_to_bev/memory_module.py:60-77
 *   hvpr_bev_fill        PointPillarScatter_Agg_Memory_1_scale.forward eval branch
 *                        pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:169-220
 *                        (PointPillarScatter.forward :14-37 when readout == NULL and scale == NULL)
 *   hvpr_build_cell_map  the index arithmetic at pointpillar_scatter.py:190-193 for externally supplied voxel_coords
 */
#ifndef HVPR_B200_H
#define HVPR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum HvprStatus {
    HVPR_OK = 0,
    HVPR_ERR_ARG = -1,          /* null pointer / negative size / inconsistent arguments */
    HVPR_ERR_UNSUPPORTED = -2,  /* configuration outside what the kernels are built for */
    HVPR_ERR_WORKSPACE = -3,    /* workspace too small */
    HVPR_ERR_CUDA = -4          /* a CUDA launch failed; cudaGetLastError text via hvpr_last_cuda_error() */
} HvprStatus;

/* Pillar grid.  lo/vs are the fp32 roundings of the YAML numbers; grid = (nx, ny, nz). */
typedef struct HvprGeom {
    float lo[3];      /* range min x,y,z */
    float vs[3];      /* voxel size x,y,z */
    int32_t grid[3];  /* nx, ny, nz */
} HvprGeom;

enum { HVPR_OVERFLOW_CONTINUE = 0, HVPR_OVERFLOW_BREAK = 1,
       HVPR_VOXELIZE_FORCE_HASH = 0x100   /* OR-ed into overflow_mode: use the open-addressing table even when the dense one fits (tests) */ };
enum { HVPR_MEM_FP32 = 0,          /* exact fp32 SIMT path */
       HVPR_MEM_BF16_RESCORE = 1   /* tcgen05 bf16 GEMM -> candidate set -> exact fp32 re-score (default) */ };

/* Folded (eval-mode BN merged into the bias-free Linear) weights of PillarVFE_Scale — HOST struct, passed by value
 * into the kernel's constant bank.  Layer sizes are the shipped cfg's (hvpr.yaml:69-75):
 * 10 -> 16 (first PFN, out//2) ; [16 | 16] -> 64 (last PFN) ; scale MLP 5 -> 16 -> 32.                         */
typedef struct HvprPfnWeights {
    float w0[16][10];   /* pfn_layers.0.linear.weight * bn scale */
    float b0[16];       /* bn shift of layer 0 (= value of a zero-padded row before ReLU) */
    float w1a[64][16];  /* pfn_layers.1: columns 0..15  (per-point activations) */
    float w1b[64][16];  /* pfn_layers.1: columns 16..31 (broadcast per-pillar max) */
    float b1[64];
    float ws0[16][5];   /* pfn_scale_layers.0 */
    float bs0[16];
    float ws1[32][16];  /* pfn_scale_layers.1 */
    float bs1[32];
} HvprPfnWeights;

/* Optional launch shape of hvpr_pfn / hvpr_bev_fill, passed with the call (NULL = defaults).  Replaces the round-1
 * process-global hvpr_tune_* knobs: two front ends in one process can no longer race on them.
 *   hvpr_pfn:      blocks_per_sm 1..3 persistent blocks of four warps per SM (0 = default 2: what shared memory admits); variant 0 = W1a tensor-core fragments in
 *                  registers (fastest alone), 1 = fragments in shared memory (fewer registers: the canvas-fill blocks of
 *                  hvpr_bev_fill fit beside the PFN blocks in the streaming schedule)
 *   hvpr_bev_fill: blocks_per_sm 0 = one 128-thread block per work item (fastest alone), 1..16 = that many persistent
 *                  blocks per SM walking the items with a grid stride; variant 0 = every canvas element is written
 *                  (feature or 0); 1 / 2 / 3 = the caller guarantees the canvases are ALL-ZERO on entry (hvpr_mem_attn's
 *                  zero_fill), only aligned runs of 32 / 64 / 128 bytes that hold a pillar are written               */
typedef struct HvprLaunchCfg {
    int32_t blocks_per_sm;
    int32_t variant;
} HvprLaunchCfg;

const char *hvpr_strerror(int status);
const char *hvpr_last_cuda_error(void);
int hvpr_version(void);
/* One-time per-process/per-device kernel attribute setup (opt-in shared memory).  Call before graph capture. */
int hvpr_init(void);

/* ---- K1 voxelize -------------------------------------------------------------------------------------------------
 * points        (n_total, pts_stride) fp32; x,y,z,intensity at columns xyz_col..xyz_col+3 (4 features are copied)
 * frame_offsets (n_frames+1) int32: frame f owns points [off[f], off[f+1])
 * max_frame_points  upper bound on any frame's point count (sizes the grid; 0 -> n_total); a frame that is longer is
 *                   truncated to its first max_frame_points points — by every kernel alike, never silently emptied
 * outputs (rows frame-major, first-seen order inside a frame, exactly as dataset.py:159-166 would collate):
 *   voxels        (>= n_frames*max_voxels rows, max_points, 4) fp32, zero-padded rows
 *   coords        (rows, 4) int32 [b, z, y, x]
 *   num_points    (rows) int32
 *   voxel_offsets (n_frames+1) int32: frame f owns rows [vo[f], vo[f+1]);  vo[n_frames] = total pillar count
 *   cell_map      (n_frames, nz*ny*nx) int32: cell -> row, -1 where empty (consumed by hvpr_bev_fill); may be NULL
 * workspace: hvpr_voxelize_workspace_bytes() bytes, contents undefined on entry.
 * Table: dense {first index, count} per cell while that is <= 4 GiB for the batch and the grid has <= 2^31 cells (every pillar grid of
 * the reference's configs); otherwise an open-addressing hash table with 64-bit cell keys (2 x max_frame_points slots per frame) —
 * then cell_map must be NULL (there is no dense cell -> row map for such grids).                                          */
size_t hvpr_voxelize_workspace_bytes(int64_t n_total, int n_frames, const HvprGeom *geom, int max_voxels);
int hvpr_voxelize(const float *points, int64_t n_total, int pts_stride, int xyz_col,
                  const int32_t *frame_offsets, int n_frames, int64_t max_frame_points,
                  const HvprGeom *geom, int max_points, int max_voxels, int overflow_mode,
                  float *voxels, int32_t *coords, int32_t *num_points, int32_t *voxel_offsets, int32_t *cell_map,
                  void *workspace, size_t workspace_bytes, void *stream);

/* frame_offsets from the leading batch-index column of a collated (n_total,5) points tensor (dataset.py:161-166). */
int hvpr_frame_offsets(const float *points5, int64_t n_total, int pts_stride, int n_frames,
                       int32_t *frame_offsets, void *stream);

/* ---- K2 fused PFN ------------------------------------------------------------------------------------------------
 * voxels (rows,max_points,4), num_points (rows), coords (rows,4) as produced above.
 * n_pillars_dev: device int32 holding the live row count (e.g. &voxel_offsets[n_frames]); NULL -> n_rows_max rows.
 * x_off/y_off/z_off: voxel/2 + range_min built by the caller with the reference's expression (pillar_vfe.py:169-171).
 * pillar_features (rows,64); scale_out (rows,32) or NULL; mask_out (rows,max_points) or NULL.
 * weights_packed: DEVICE image made by hvpr_pfn_pack from the same weights_host (hvpr_pfn_packed_bytes() bytes), or NULL
 * (every block then rebuilds its tensor-core weight fragments from the by-value weights: ~10 % slower).              */
size_t hvpr_pfn_packed_bytes(void);
int hvpr_pfn_pack(const HvprPfnWeights *weights_host, void *weights_packed, void *stream);
int hvpr_pfn(const float *voxels, const int32_t *num_points, const int32_t *coords,
             const int32_t *n_pillars_dev, int64_t n_rows_max, int max_points,
             const HvprPfnWeights *weights_host, const HvprGeom *geom, float x_off, float y_off, float z_off,
             float *pillar_features, float *scale_out, float *mask_out, const void *weights_packed,
             const HvprLaunchCfg *launch, void *stream);

/* ---- K3 memory attention -----------------------------------------------------------------------------------------
 * pillars (rows,64) fp32; mem_weight (M,64) fp32; readout (rows,64) fp32.
 * mem_weight_bf16: (M_pad,64) bf16 copy made by hvpr_mem_pack_bf16 (M_pad = M rounded up to 256); required for
 *                  HVPR_MEM_BF16_RESCORE, ignored for HVPR_MEM_FP32.
 * topk_idx_out: optional (rows,k) int32 — the selected memory items (unordered set), for tests.
 * zero_fill: optional side job (NULL = none): in stream order, every listed range is all-zero when the call completes.
 *            The tcgen05 path is compute-bound and leaves HBM idle, so its TMA-producer thread streams the zeros out
 *            with bulk async stores while the tiles are processed (the BEV canvases of hvpr_bev_fill's
 *            canvas-is-zero mode); measured, the write stream slows the latency-bound kernel by more than the fill
 *            saves, so nothing in the shipped pipeline uses it; the fp32 path and the no-rows case
 *            use cudaMemsetAsync.  The ranges must not overlap anything the call reads or writes.          */
typedef struct HvprZeroFill {
    void *ptr[4];       /* device ranges, 16-byte aligned */
    uint64_t bytes[4];  /* multiples of 16; 0 = unused slot */
    int32_t n;          /* ranges in use, <= 4 */
} HvprZeroFill;
size_t hvpr_mem_attn_workspace_bytes(int64_t n_rows_max, int M, int precision_mode);
int hvpr_mem_pack_bf16(const float *mem_weight, int M, int C, void *mem_weight_bf16, void *stream);
int hvpr_mem_attn(const float *pillars, const int32_t *n_pillars_dev, int64_t n_rows_max,
                  const float *mem_weight, const void *mem_weight_bf16, int M, int C, int k, int precision_mode,
                  float *readout, int32_t *topk_idx_out, void *workspace, size_t workspace_bytes,
                  const HvprZeroFill *zero_fill, void *stream);

/* ---- K4 BEV canvas gather-fill -----------------------------------------------------------------------------------
 * Writes every canvas element exactly once (feature or 0), NCHW, x fastest (pointpillar_scatter.py:192,217-218):
 *   spatial       (n_frames, ca+cb, ny, nx)  channels [feat_a (pillar) | feat_b (memory readout)]
 *   spatial_scale (n_frames, cs, ny, nx)     or NULL when cs == 0
 * cell_map (n_frames, ny*nx) int32: cell -> pillar row or -1.                                                        */
int hvpr_bev_fill(const float *feat_a, int ca, const float *feat_b, int cb, const float *feat_s, int cs,
                  const int32_t *cell_map, int n_frames, int nx, int ny,
                  float *spatial, float *spatial_scale, const HvprLaunchCfg *launch, void *stream);

/* cell_map from externally supplied coords (rows,4) int32 [b,z,y,x] (module API fed by a foreign voxelizer).         */
int hvpr_build_cell_map(const int32_t *coords, const int32_t *n_pillars_dev, int64_t n_rows_max,
                        int n_frames, int nx, int ny, int32_t *cell_map, void *stream);

/* ==== N4: training-branch hybrid aggregation, FORWARD only ============================================
 * get_score (pointpillar_scatter.py:67-83: softmax over the frame's points, top-k points per pillar, exact re-score, softmax_k,
 * readout) is hvpr_mem_attn with the frame's point features (np, 64) as `mem_weight`, M = np and HVPR_MEM_FP32 (exact, any M);
 * topk_idx_out then names the k positive points of every pillar.
 * hvpr_mem_train_forward = MemoryUnit_Agg.forward training branch (memory_module.py:31-59, hard_shrink_relu :85-87):
 *   pillars (nv,64); points_positive (nv*k,64) the k positive point features of every pillar, pillar-major;
 *   memory_positive_ws (nv*k,64) scratch that receives `memory_positive` (:49-50); output (nv,64).  M <= 2048.
 * hvpr_mse_loss = get_mem_loss (anchor_head_template.py:262-275): loss_out[0] = scale * sum((a-b)^2) over n elements, with
 *   scale = mem_weight / (n * rows) set by the caller; partials_ws: 1024 floats.  Deterministic (no atomics).                 */
int hvpr_mem_train_forward(const float *pillars, int64_t nv, const float *points_positive, const float *mem_weight,
                           int M, int C, int k, float shrink_thres, float *memory_positive_ws, float *output, void *stream);
int hvpr_mse_loss(const float *a, const float *b, int64_t n, double scale, float *partials_ws, float *loss_out, void *stream);

/* ==== N1 (first "next" row): BaseBEVBackbone_Scale convolutions ======================================
 * Replaces the nn.Conv2d / nn.ConvTranspose2d + BatchNorm2d + ReLU stacks built at
 * pcdet/models/backbones_2d/base_bev_backbone.py:150-213 and run by the eval forward at :280-315, and the
 * SpatialAttention gate of pcdet/models/backbones_2d/spatial_attention.py:47-63.
 * Activations are NHWC bf16 with a channel stride (in_cs / out_cs elements per pixel); eval-mode BN is folded into the
 * weights (scale) and the bias (shift) by the caller.                                                               */
typedef struct HvprConvArgs {
    const void *in;          /* (n, h_in, w_in, in_cs) bf16 */
    int32_t n, h_in, w_in, in_cs;
    int32_t c_in;            /* channels contracted: multiple of 64 (pad with zero channels / zero weights) */
    int32_t ksize, stride;   /* 3 (zero padding 1) with stride 1 | 2 (even h_in, w_in), or 1 with stride 1 */
    const void *w_packed;    /* hvpr_conv_pack_weights image */
    int32_t n_total;         /* GEMM columns: c_out (out_mode 0) or up*up*c_out (out_mode 1), <= 2048 */
    int32_t bn;              /* column tile: 32, 64, 128 or 256, dividing n_total (same value as at pack time) */
    const float *bias;       /* (n_total) fp32 or NULL */
    int32_t relu;
    const float *gate;       /* out_mode 0: optional (n, h_out, w_out) fp32 multiplier applied after the ReLU */
    const void *residual;    /* out_mode 0: optional (n, h_out, w_out, res_cs) bf16 added after the gate */
    int32_t res_cs;
    int32_t out_mode;        /* 0: bf16 NHWC (n, h_out, w_out, out_cs), channels [out_c_off, out_c_off + n_total)
                                1: ConvTranspose2d(k = up, stride = up) pixel shuffle into fp32 NCHW
                                   (n, out_ctot, h_out*up, w_out*up), channels [out_c_off, out_c_off + c_out);
                                   GEMM column = (dy*c_out + co)*up + dx
                                2: fp32 NHWC (n, h_out, w_out, out_cs), channels [out_c_off, out_c_off + n_total) (no gate / residual)
                                3: ConvTranspose2d(k = up, stride = up) pixel shuffle into bf16 NHWC
                                   (n, h_out*up, w_out*up, out_cs), channels [out_c_off, out_c_off + c_out);
                                   GEMM column = (dy*up + dx)*c_out + co                                          */
    void *out;
    int32_t out_cs, out_c_off;
    int32_t up, c_out, out_ctot;
} HvprConvArgs;

/* w_ntc: (n_total, taps, c_in) fp32 DEVICE, taps row-major (dy, dx).  out_packed: hvpr_conv_packed_bytes() bytes. */
size_t hvpr_conv_packed_bytes(int n_total, int taps, int c_in);
int hvpr_conv_pack_weights(const float *w_ntc, int n_total, int taps, int c_in, int bn, void *out_packed, void *stream);
/* args is a HOST struct; TMA tensor maps are encoded on the host per call (no device sync, graph-capturable). */
int hvpr_conv2d(const HvprConvArgs *args, void *stream);
/* K4 in the layout the backbone consumes: channels-last bf16 canvases, every element written once (feature or 0).
 *   spatial_nhwc (n_frames, ny, nx, spatial_cs) bf16: channels [feat_a | feat_b | 0...]
 *   scale_nhwc   (n_frames, ny, nx, scale_cs)   bf16: channels [feat_s | 0...]      (pointpillar_scatter.py:204-218)   */
int hvpr_bev_fill_nhwc_bf16(const float *feat_a, int ca, const float *feat_b, int cb, const float *feat_s, int cs,
                            const int32_t *cell_map, int n_frames, int nx, int ny,
                            void *spatial_nhwc, int spatial_cs, void *scale_nhwc, int scale_cs, void *stream);
/* fp32 NCHW (n,c,h,w) -> bf16 NHWC (n,h,w,out_cs), channels [0,c); other channels of out are left untouched. */
int hvpr_nchw_to_nhwc_bf16(const float *in, int n, int c, int h, int w, void *out, int out_cs, void *stream);
/* gate = sigmoid(BN(conv3x3_{2->1}([max_c y, mean_c y]) + b)); w18_host = folded weights [(ch*3+dy)*3+dx] (HOST),
 * bias = folded scalar; pooled_ws: (n*h*w*2) fp32 scratch; gate_out: (n,h,w) fp32.                                 */
int hvpr_attention_gate(const void *y_nhwc_bf16, int n, int h, int w, int cs, int c, const float *w18_host, float bias,
                        float *pooled_ws, float *gate_out, void *stream);

/* ==== N2: AnchorHeadSingle eval (anchor_head_single.py:109-145) ======================================================
 * The three 1x1 head convolutions run as one hvpr_conv2d (out_mode 2) whose fp32 NHWC rows hold
 * [cls A*C | box A*7 | dir A*bins | pad] at channel offsets cls_off / box_off / dir_off (dir_off < 0: no direction classifier).
 * hvpr_head_decode = generate_predicted_boxes (anchor_head_template.py:293-340): ResidualCoder.decode_torch against
 * anchors (h*w*A, 7) [x,y,z,dx,dy,dz,r] and the direction fix-up rg = limit_period(rg - dir_offset, dir_limit_offset,
 * 2*pi/bins) + dir_offset + 2*pi/bins * argmax(dir).  cls_out (n, h*w*A, C) raw logits; box_out (n, h*w*A, 7).       */
int hvpr_head_decode(const float *head_nhwc, int n, int h, int w, int cs, int A, int C, int cls_off, int box_off,
                     int dir_off, int num_dir_bins, const float *anchors, float dir_offset, float dir_limit_offset,
                     float *cls_out, float *box_out, void *stream);

/* ==== N3: single-stage post-processing ================================================================================
 * Detector3DTemplate.post_processing, class-agnostic branch (pcdet/models/detectors/detector3d_template.py:168-260) +
 * class_agnostic_nms (pcdet/models/model_utils/model_nms_utils.py:6-25): sigmoid / max over classes, score threshold,
 * top nms_pre_max (<= 4096) by score, greedy rotated-BEV-IoU NMS, first nms_post_max survivors per frame.
 * The reference's IoU/NMS op (`iou3d_nms`, setup.py:53-62) is not in the tree: parity for it is UNPINNED.
 * cls_preds (n_frames, n_boxes, num_class), box_preds (n_frames, n_boxes, 7) [x,y,z,dx,dy,dz,heading] fp32 device.
 * outputs (capacity nms_post_max per frame, in descending score order): out_boxes (n_frames, post, 7), out_scores,
 * out_labels (1-based), out_index (anchor index of each kept box), out_count (n_frames).                             */
size_t hvpr_post_process_workspace_bytes(int n_frames, int64_t n_boxes);
int hvpr_post_process(const float *cls_preds, const float *box_preds, int n_frames, int64_t n_boxes, int num_class,
                      int cls_normalized, float score_thresh, int nms_pre_max, int nms_post_max, float nms_thresh,
                      float *out_boxes, float *out_scores, int32_t *out_labels, int32_t *out_index, int32_t *out_count,
                      void *workspace, size_t workspace_bytes, void *stream);
/* Pairwise 3-D IoU (n, m) of boxes [x, y, z, dx, dy, dz, heading] (z = centre): the boxes_iou3d_gpu of the reference's absent iou3d_nms
 * op, used by generate_recall_record (detector3d_template.py:277-310).                                                              */
int hvpr_boxes_iou3d(const float *boxes_a, int n, const float *boxes_b, int m, float *iou, void *stream);


#ifdef __cplusplus
}
#endif
#endif /* HVPR_B200_H */
