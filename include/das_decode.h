/*
 * das_decode.h -- C ABI of the B200-native DAS dense-head inference decode.
 *
 * The reference (wangzt-halo/das, an mmdetection3d fork) has NO native interface for this path:
 * it is eager PyTorch + NumPy.  The entry points below are what a binding for the
 * path replaces, one reference function (file:line, relative to the reference root) per launcher:
 *
 *   das_score_topk ............ DASHead._get_poses_single, das_head.py:708-723
 *                               (two sigmoids, score product, per-level topk(nms_pre))
 *   das_gather_refine_assemble  das_head.py:720-749 (gather + joint assembly), das_head.py:237-262
 *                               (eval tail) and the LAST RecursiveUpdateLayer evaluated sparsely at
 *                               the selected centres (recursive_update.py:186-197, 9-82)
 *   das_refine_dense_layer .... one full-map RecursiveUpdateLayer, recursive_update.py:220-235
 *                               (only needed for layers 1..L-1 when num_layers > 1)
 *   das_nms_backproject ....... das_head.py:751-794 (score_thr, areas, OKS-NMS, nms_post),
 *                               pose_nms.py:51-126, cmupanoptic_mono_dataset.py:391-401 (depth
 *                               de-normalisation) and mytools/vis_3d.py:16-26 (pixel2world)
 *   das_plan_* ................ DASHead.get_poses, das_head.py:653-688 (the whole call)
 *
 * Conventions: every function returns DAS_OK (0) or a negative das_status; launchers take raw
 * DEVICE pointers plus explicit sizes and a cudaStream_t (passed as void*), never allocate, never
 * synchronise, and are CUDA-graph capturable.  Only das_plan_create / das_plan_destroy allocate.
 * All tensors are fp32 unless stated; "NCHW" = [B, C, H, W] contiguous, "NHWC" = [B, H, W, C].
 */
#ifndef DAS_DECODE_H_
#define DAS_DECODE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAS_MAX_LEVELS 5      /* reference pyramids have 4 (exp_panoptic.py:35) or 5 (_base_/models/das.py:30) */
#define DAS_MAX_LAYERS 4      /* recursive_update.num_layers: 1 (Panoptic), 2 (MuPoTS), 3 (BASELINE config 3) */
#define DAS_MAX_JOINTS 32
#define DAS_MAX_NMS_PRE 2048  /* per-level top-k capacity (reference configs use 1000) */
#define DAS_CAM_DOUBLES 18    /* K[0,:3], K[1,:3], R row-major 3x3, t[3] */
#define DAS_MAX_PEERS 15      /* other GPUs of one NVSwitch box that receive a copy of a rank's result block */
#define DAS_NUM_STAGES 5      /* score_topk | dense layers | refine phases 1-2 (tensor-core mode) | refine GEMM + finish / SIMT refine | nms+backproject */

typedef enum das_status {
    DAS_OK = 0,
    DAS_ERR_ARG = -1,          /* bad argument (null pointer, size out of range) */
    DAS_ERR_CUDA = -2,         /* a CUDA runtime call failed; see das_last_error() */
    DAS_ERR_CAPACITY = -3,     /* nms_pre / candidate count / joints beyond compiled capacity */
    DAS_ERR_UNSUPPORTED = -4   /* feature not built (e.g. channel count other than 128/256/512) */
} das_status;

/* Element type of the head-output maps cls / ctr / pose (das_levels.in_dtype).  The reference's shipped fp16 mode
 * (das_head.py:180,218 out_fp16=True, exp_panoptic.py:222) hands get_poses fp16 tensors; the kernels read them in
 * place (half the scan / gather bytes) and compute in fp32 -- bit-identical to up-casting the maps first. */
#define DAS_DTYPE_F32 0
#define DAS_DTYPE_F16 1
#define DAS_DTYPE_BF16 2

/* One pyramid level.  All pointers are device pointers owned by the caller.  cls / ctr / pose point at elements of
 * das_levels.in_dtype (declared const float* for the common fp32 case; cast for fp16 / bf16). */
typedef struct das_level_desc {
    const float* cls;                    /* [B,1,H,W] centre-heat-map logits (cls_score)          */
    const float* ctr;                    /* [B,1,H,W] centerness logits                           */
    const float* pose;                   /* [B,3+6J,H,W] NCHW. refine=1: RAW predictor output     */
                                         /*                    refine=0: final pose_pred (ref. API) */
    const float* feats[DAS_MAX_LAYERS];  /* refine=1: per refinement layer, NHWC [B,H,W,C]        */
    int32_t H, W, stride;
    float scale_offset, scale_depth, scale_uv, scale_d;   /* mmcv Scale params, das_head.py:171-173 */
} das_level_desc;

typedef struct das_levels {
    int32_t n_levels;
    int32_t batch;
    int32_t in_dtype;                    /* DAS_DTYPE_* of cls / ctr / pose of every level (feats are always fp32) */
    int32_t reserved_;
    das_level_desc lv[DAS_MAX_LEVELS];
} das_levels;

/* Head + test_cfg keys of the path (configs/_base_/models/das.py:24-51, exp_panoptic.py:31-53). */
typedef struct das_decode_cfg {
    int32_t num_joints;      /* J */
    int32_t root_idx;
    int32_t num_heads;       /* recursive_update.num_heads (4) */
    int32_t feat_channels;   /* recursive_update.feat_channels (256) */
    int32_t num_layers;      /* recursive_update.num_layers */
    float depth_factor;
    float z_norm;
    int32_t nms_pre;         /* <=0: keep every cell (das_head.py:716-717) */
    int32_t nms_post;        /* <=0: skip OKS-NMS (das_head.py:770-772) */
    float nms_thr;
    float score_thr;         /* <=0: no threshold (das_head.py:763-764) */
    int32_t peak_kernel;     /* 0/1: reference behaviour; 3: north-star 3x3 max-pool peak mask */
    int32_t refine;          /* 1: pose maps are raw, run refinement + eval tail; 0: maps are final */
    double dataset_depth_factor; /* cmupanoptic_mono_dataset.py:399 (dataset-side factor, 1.0) */
    int32_t nms_soft;        /* test_cfg.nms_type != 'hard': soft_oks_nms (pose_nms.py:129-194, das_head.py:789-790) */
    int32_t reserved_;
} das_decode_cfg;

/* Device buffers of one decode.  CT = candidate slots per image (das_candidate_slots()),
 * P = das_output_slots(). Slot order is level-major, rank order inside a level. */
typedef struct das_buffers {
    /* stage 1-2 */
    float*   cand_score;   /* [B,CT]  sigmoid(cls)*sigmoid(ctr) of the selected cell              */
    int32_t* cand_index;   /* [B,CT]  flat cell index y*W+x inside its level                      */
    /* stage 3-4 */
    float*   cand_pose;    /* [B,CT,J,3] joints: x,y original-image px, z = d + z_root*sqrt(sx*sy) */
    float*   cand_center;  /* [B,CT,3]                                                            */
    /* stage 5 outputs */
    int32_t* out_count;    /* [B]                                                                 */
    float*   out_score;    /* [B,P]                                                               */
    int32_t* out_slot;     /* [B,P]   candidate slot each output row came from                    */
    float*   out_pose;     /* [B,P,J,3]                                                           */
    float*   out_center;   /* [B,P,3]                                                             */
    double*  out_cam;      /* [B,P,J,3] camera space (vis_3d.py x2)                               */
    double*  out_world;    /* [B,P,J,3] world space  (vis_3d.py x3)                               */
} das_buffers;

/* Peer copies of a rank's packed output block (the one collective of the path).  delta[q] is the BYTE
 * distance from the rank's local output block to its copy inside peer q's gathered buffer (a peer-mapped device
 * address: cudaIpcOpenMemHandle / cudaDeviceEnablePeerAccess); das_nms_backproject_peers stores every result value
 * locally AND at +delta[q] for q < n, over NVLink.  seq / ticket (both inside / beside the LOCAL block, device int32,
 * or NULL): the last CTA increments *seq locally and on every peer after a system-scope fence, so a consumer on a
 * peer that observes seq == s may read step s. */
typedef struct das_peer_blocks {
    int32_t n;
    int32_t reserved_;
    int64_t delta[DAS_MAX_PEERS];
    int32_t* seq;
    int32_t* ticket;
} das_peer_blocks;

const char* das_version(void);
/* sizeof of the structs that cross the ABI by value or pointer: {das_levels, das_decode_cfg, das_buffers, das_row_cache,
 * das_refine_scratch, das_peer_blocks}.  A binding checks these against its own mirrors at load time (a stale library
 * must fail loudly, not corrupt memory). */
void das_abi_struct_sizes(int32_t out[6]);
const char* das_last_error(void);

/* slot bookkeeping (host-side pure functions) */
int32_t das_level_slots(int32_t H, int32_t W, int32_t nms_pre);          /* min(HW, nms_pre) rule */
int32_t das_candidate_slots(const das_levels* lv, int32_t nms_pre);      /* CT = sum over levels  */
int32_t das_output_slots(int32_t cand_slots, int32_t nms_post);          /* P                     */

/* ---- stage launchers (device pointers, stream, no allocation, no sync) ------------------------ */

/* Stage 1+2: fused sigmoid*sigmoid (+ optional 3x3 peak mask) and exact per-level top-k, ties to
 * the lower cell index.  d_levels: device copy of das_levels.  scratch: >= B*sum(HW) uint32. */
int das_score_topk(const das_levels* d_levels, const das_levels* h_levels, int32_t nms_pre,
                   int32_t peak_kernel, float* cand_score, int32_t* cand_index, int32_t cand_slots,
                   uint32_t* scratch, void* stream);

/* Stage 3+4(+eval tail): gather at the selected cells, sparse last-layer refinement, assembly.
 * weights: das_pack_weights() layout of the LAST layer; prev_uvd: NULL (use scaled raw uvd of
 * lv.pose) or [n_levels] device pointers to joint-major [B][J][HW][4] maps produced by dense layers.
 * scale_xy: [B,2] device (img_metas['scale_factor'][:2]). */
int das_gather_refine_assemble(const das_levels* d_levels, const das_levels* h_levels,
                               const das_decode_cfg* cfg, const float* weights,
                               const float* const* prev_uvd, const float* scale_xy,
                               const float* cand_score, const int32_t* cand_index, int32_t cand_slots,
                               float* cand_pose, float* cand_center, int32_t* work_counter,
                               void* stream);

/* Device row cache of the host zero-copy mode; see das_refine_row_cache below. */
typedef struct das_row_cache {
    void* table;        /* das_row_cache_table_bytes(table_bits) bytes of device memory */
    float* rows;        /* [max_rows][feat_channels] device */
    float* cand_rows;   /* [B*CT][feat_channels] device, or NULL */
    int32_t table_bits;
    int32_t max_rows;
} das_row_cache;

/* Scratch of the tensor-core sparse refinement (device memory, caller- or plan-owned).  row_cap >= B*CT*32. */
typedef struct das_refine_scratch {
    float* unique_rows;    /* [J][row_cap][8]   distinct sampled cells of every joint: {feature-row pointer (2 words), previous
                                                offset u, v, d at that cell, -, -, -}                                          */
    float* unique_out;     /* [J][row_cap][8]   per distinct cell: {blended O u, v, d, confidence u, v, d, -, -}                */
    float* row_records;    /* [B*CT*J][32][4]   per (item, head, corner): {index into the joint's distinct-row list (int32 bits,
                                                -1 = outside the map), bilinear weight, head offset x, y}                      */
    float* item_records;   /* [B*CT*J][8]       assembly record {Px, Py, zq, sx, sy, stride, -, -}                             */
    int32_t* valid_list;   /* [B*CT]            candidates above score_thr (b*CT + slot)                                        */
    int32_t* counters;     /* [4 + DAS_MAX_JOINTS] [0] work queue, [1] number of valid candidates, [2] reserved (peer-store
                                                ticket of the plan), [4+j] distinct rows of joint j                            */
    int32_t row_cap;
    int32_t reserved_;
    const float* const* prev_planes; /* NULL, or device array [n_levels] of the projection planes das_dense_project_tc wrote for
                                        layer L-2 (num_layers > 1): das_refine_heads then evaluates that layer's progressive
                                        sampling on demand, only at the cells the last layer looks at (<= 33 per item), and
                                        ignores prev_uvd -- layer L-2 needs no das_refine_dense_layer sampling pass            */
} das_refine_scratch;

/* Tensor-core variant of stage 3+4 (feat_channels = 256, num_heads = 4), three launches:
 *   das_refine_heads   phases 1-2 per (candidate, joint) item: the sampling position of each of the 32 (head, corner) rows;
 *                      the item's DISTINCT sampled cells are appended to the joint's row list (the gate / value / confidence
 *                      projections depend on (cell, joint) only, and the 8 heads x 4 corners mostly land on a handful of
 *                      cells), the 32 row records point into it; assembly record; the centre (joint 0); valid_list
 *   das_refine_tc      the distinct rows of every joint as a gathered tcgen05 GEMM in 128-row tiles (split=1: 3xTF32,
 *                      fp32-level accuracy; split=0: one TF32 pass) + bias / gate / blend epilogue -> unique_out
 *   das_refine_finish  per item: bilinear weights, corner sums, softmax over the 2*nh heads, eval tail, assembly -> cand_pose
 * panels: das_pack_tc_panels() image of the last layer's packed weights (das_tc_panel_bytes() bytes). */
int das_refine_heads(const das_levels* d_levels, const das_levels* h_levels, const das_decode_cfg* cfg,
                     const float* weights, const float* const* prev_uvd, const float* scale_xy,
                     const float* cand_score, const int32_t* cand_index, int32_t cand_slots,
                     const das_refine_scratch* scratch, float* cand_center, const das_row_cache* rc /* NULL = off */,
                     void* stream);
int das_refine_tc(const das_levels* d_levels, const das_levels* h_levels, const das_decode_cfg* cfg,
                  const float* weights, const void* panels, int32_t cand_slots,
                  const das_refine_scratch* scratch, int32_t split, void* stream);
int das_refine_finish(const das_levels* h_levels, const das_decode_cfg* cfg, int32_t cand_slots,
                      const das_refine_scratch* scratch, float* cand_pose, void* stream);
int das_pack_tc_panels(const das_decode_cfg* cfg, const float* packed_weights, void* panels, void* stream);
/* Row cache for the host zero-copy mode (feature maps read in place from pinned HOST memory).  On device memory L2
 * absorbs the ~10x re-use of feature rows between heads / joints / candidates; reads of host memory are not
 * deduplicated that way, so these passes make every distinct row cross PCIe once:
 *   das_row_cache_clear    empties the table (before das_refine_cand_rows / das_refine_heads of a batch)
 *   das_refine_cand_rows   copies F(p) of every candidate above score_thr into rc->cand_rows (read by all J joints)
 *   das_refine_heads(rc)   reads F(p) from rc->cand_rows and leaves the target-corner rows it fetched in the cache
 *   das_refine_row_cache   between das_refine_heads and das_refine_tc: copies every still-missing row of the joints'
 *                          distinct-row lists into rc->rows and re-points the list entries at the copies.
 * A full row buffer leaves the remaining records pointing at the host.  2^table_bits should be at least twice the
 * number of row records (B*CT*J*32). */
int64_t das_row_cache_table_bytes(int32_t table_bits);
int das_row_cache_clear(const das_row_cache* rc, void* stream);
int das_refine_cand_rows(const das_levels* d_levels, const das_levels* h_levels, const das_decode_cfg* cfg,
                         const float* cand_score, const int32_t* cand_index, int32_t cand_slots,
                         const das_row_cache* rc, void* stream);
int das_refine_row_cache(const das_decode_cfg* cfg, const das_refine_scratch* scratch, const das_row_cache* rc, void* stream);
int64_t das_tc_panel_bytes(const das_decode_cfg* cfg);
/* profiling aid: per-CTA cycle counters of das_refine_tc's warp roles ([148][16] int64 device buffer; NULL = off) */
int das_tc_set_debug_buffer(long long* dev_buf);
/* same for das_dense_project_tc: [148][8] int64 {mma: wait acc_free, wait a_full, issue | producer group 0: wait TMA, wait TMEM
 * slot, work | epilogue group 0: wait acc_full | CTA total}; the caller zeroes the buffer; NULL = off */
int das_dense_set_debug_buffer(long long* dev_buf);

/* One dense refinement layer over a whole level (layers 1..L-1 when num_layers > 1): 1x1 projection + gated
 * blend into proj (scratch [B][J][HW][16] fp32), then the progressive sampling into uvd_out (joint-major
 * [B][J][HW][4] = u, v, d, pad).  uvd_in NULL -> scaled raw uvd from lv.pose.  tc_panels: das_pack_dense_panels() image of this layer's packed
 * weights -> the projection runs on the tensor cores (tcgen05 3xTF32; feat_channels = 256, num_heads = 4);
 * NULL -> fp32 SIMT projection. */
int das_refine_dense_layer(const das_levels* d_levels, const das_levels* h_levels, int32_t level,
                           int32_t layer, const das_decode_cfg* cfg, const float* weights,
                           const void* tc_panels, const float* uvd_in, float* uvd_out, float* proj, void* stream);
int das_dense_project_tc(const das_levels* d_levels, const das_levels* h_levels, int32_t level, int32_t layer,
                         const das_decode_cfg* cfg, const float* weights, const void* panels,
                         const float* uvd_in, float* proj, void* stream);
int das_pack_dense_panels(const das_decode_cfg* cfg, const float* packed_weights, void* panels, void* stream);
int64_t das_dense_panel_bytes(const das_decode_cfg* cfg);

/* Stage 5: score_thr, OKS-NMS, nms_post, output packing, depth de-norm + back-projection.
 * cam: [B,DAS_CAM_DOUBLES] device doubles. */
int das_nms_backproject(const das_decode_cfg* cfg, int32_t batch, int32_t cand_slots,
                        const float* cand_score, const float* cand_pose, const float* cand_center,
                        const double* cam, das_buffers out, void* stream);

/* Same, with the result all-gather fused in: every output value is also stored into each peer's copy of this rank's
 * block (see das_peer_blocks). peers == NULL or peers->n == 0: identical to das_nms_backproject. */
int das_nms_backproject_peers(const das_decode_cfg* cfg, int32_t batch, int32_t cand_slots,
                              const float* cand_score, const float* cand_pose, const float* cand_center,
                              const double* cam, das_buffers out, const das_peer_blocks* peers, void* stream);

/* The same collective as its own small kernel behind das_nms_backproject (what das_plan enqueues): a few CTAs copy
 * `bytes` (multiple of 16, 16-byte aligned) of the rank's packed block to every peer and publish the sequence word. */
int das_peer_publish(const das_peer_blocks* peers, const void* local_block, int64_t bytes, void* stream);

/* Repack the four 1x1 convolutions of one RecursiveUpdateLayer (nn.Conv2d weight [O,C] + bias [O],
 * recursive_update.py:171-180) into the joint-major layout the kernels read:
 * dst[j][17][C] rows = {sampling_offset 2*nh, update_weight 3, update_offset_value 3, sampling_conf 3}
 * followed by dst_bias[j][17].  All pointers device; dst holds J*(2nh+9)*(C+1) floats. */
int das_pack_weights(const das_decode_cfg* cfg, const float* so_w, const float* so_b,
                     const float* sc_w, const float* sc_b, const float* uw_w, const float* uw_b,
                     const float* uv_w, const float* uv_b, float* dst, void* stream);
int64_t das_packed_weight_floats(const das_decode_cfg* cfg);

/* ---- plan: the whole get_poses call, CUDA-graph captured --------------------------------------- */
typedef struct das_plan das_plan;

int das_plan_create(const das_decode_cfg* cfg, const das_levels* shape /* H,W,stride,batch only */,
                    das_plan** out);
void das_plan_destroy(das_plan* plan);
/* device weights of layer `layer` (nn.Conv2d layouts); repacked on `stream`. */
int das_plan_set_weights(das_plan* plan, int32_t layer, const float* so_w, const float* so_b,
                         const float* sc_w, const float* sc_b, const float* uw_w, const float* uw_b,
                         const float* uv_w, const float* uv_b, void* stream);
/* bind device inputs (pointers + per-level Scale values); cheap, no graph rebuild. */
int das_plan_bind(das_plan* plan, const das_levels* levels, void* stream);
/* per-image metas from HOST memory: scale_xy [B,2] fp32, cam [B,18] fp64. */
int das_plan_set_metas(das_plan* plan, const float* scale_xy, const double* cam, void* stream);
/* enqueue one decode on `stream`. mode 0: eager launches; 1: CUDA-graph replay (captured on first
 * use); 2: graph replay with event-record nodes at the stage boundaries (for das_plan_stage_ms). */
int das_plan_run(das_plan* plan, void* stream, int32_t mode);
/* per-stage milliseconds of the last mode-2 replay (stream must be synchronised): ms[DAS_NUM_STAGES] */
int das_plan_stage_ms(das_plan* plan, float* ms);
/* all out_* buffers are carved from one device block (one D2H, or one NCCL all-gather across ranks) */
int das_plan_output_block(const das_plan* plan, void** ptr, int64_t* bytes);
/* refinement implementation: 0 = fp32 SIMT, 1 = tensor cores 3xTF32 (default when feat_channels = 256 and
 * num_heads = 4), 2 = tensor cores, single TF32 pass (looser accuracy). Call before the first run. */
int das_plan_set_refine_mode(das_plan* plan, int32_t mode);
/* Programmatic dependent launch along the decode's kernel chain: 1 = every kernel may be scheduled while its predecessor
 * still runs and waits on the SM for it (hides launch gaps and ramps: the latency mode, for one decode at a time),
 * 0 = plain stream order (the throughput mode, for several independent decodes in flight on different streams: waiting
 * CTAs would take SM resources from them), -1 = auto (default): on when batch * candidate slots * joints <= 24 * 148. */
int das_plan_set_pdl(das_plan* plan, int32_t mode);
/* num_layers > 1 on the tensor-core path: 1 (default) = layer L-2 runs its projection only and the sparse last layer
 * evaluates that layer's progressive sampling on demand, at the <= 33 cells per (candidate, joint) it looks at
 * (das_refine_scratch.prev_planes); 0 = every dense layer samples its whole map like recursive_update.py:220-235.  Same
 * values either way (one device function); layers 0..L-3 are always dense because the next projection blends with their
 * output at every cell. */
int das_plan_set_on_demand_sampling(das_plan* plan, int32_t on);
int das_plan_buffers(const das_plan* plan, das_buffers* out, int32_t* cand_slots, int32_t* out_slots);
/* Caller-owned output block: the plan writes its out_* buffers into `block` (device memory, 256-B aligned, at least
 * das_plan_output_block() bytes + 256 for the sequence word) instead of its own allocation -- e.g. a slice of one
 * contiguous staging / gathered buffer shared by several plans, so collecting results needs no copy.  Any captured
 * graph is dropped and re-captured by the next run; das_plan_buffers / das_plan_output_block report the new pointers. */
int das_plan_set_output_block(das_plan* plan, void* block, int64_t bytes);
/* Fused result all-gather over NVLink: peer_blocks[q] = address (mapped into THIS process) of this rank's slot inside
 * peer q's gathered buffer, same layout as the local block.  n_peers == 0 switches it off.  With it on, the local
 * block carries a sequence word right behind the packed outputs (offset das_plan_output_block() bytes, int32). */
int das_plan_set_peer_blocks(das_plan* plan, int32_t n_peers, void* const* peer_blocks);
/* With peer blocks set, das_plan_run enqueues the publication (das_peer_publish) on a side stream of the plan behind the
 * decode, so `stream` is free for the next decode while the NVLink stores are in flight; the plan's own next run waits for
 * it before overwriting the block.  das_plan_publish_wait makes `stream` wait for the last publication (no-op without
 * peers). */
int das_plan_publish_wait(das_plan* plan, void* stream);
/* Device memory that other processes on the box can map (cudaIpc*): das_ipc_alloc on the owner, the 64-byte handle is
 * sent to the peers through any host channel, das_ipc_open on each peer (enables peer access), das_ipc_close there,
 * das_ipc_free on the owner. */
int das_ipc_alloc(int64_t bytes, void** dev_ptr, unsigned char handle[64]);
int das_ipc_open(const unsigned char handle[64], void** dev_ptr);
int das_ipc_close(void* dev_ptr);
int das_ipc_free(void* dev_ptr);
int64_t das_plan_kernel_launches(const das_plan* plan);   /* kernels enqueued by das_plan_run so far */

/* Host-buffer entry (the end-to-end call): copies every input from HOST memory (pinned for full
 * speed) to plan-owned device staging, runs the decode and copies the packed results back to the
 * host arrays of `host_out` (same layout as das_buffers, any pointer may be NULL), then
 * synchronises `stream`. `levels` holds HOST pointers here. */
int das_plan_run_host(das_plan* plan, const das_levels* levels, const float* scale_xy,
                      const double* cam, das_buffers host_out, void* stream);
/* The same call without the final synchronisation: everything (H2D copies, decode, D2H copies) is enqueued on `stream` and
 * the call returns; the caller synchronises the stream (or an event recorded behind the call) before it reads `host_out`,
 * and does not touch the input maps, `host_out` or this plan again before that.  Two plans on two streams driven in turn
 * this way keep PCIe busy across calls: batch n+1 is copied while the kernels of batch n read their rows in place and
 * while its results travel back (bench.py `e2e`).  scale_xy / cam are consumed before the call returns. */
int das_plan_run_host_async(das_plan* plan, const das_levels* levels, const float* scale_xy,
                            const double* cam, das_buffers host_out, void* stream);
int64_t das_plan_h2d_bytes(const das_plan* plan);
/* das_plan_run_host transfer policy: 0 = bulk H2D copy of every input map (default); 1 = only the logit planes are
 * copied, the pose and feature maps (of which the decode touches ~5 %) are read in place from PINNED host memory
 * by the gather kernels (falls back to 0 for pageable memory); 2 = as 1, plus das_refine_row_cache in front of the
 * tensor-core sampling phase so that every distinct feature row crosses PCIe once. */
int das_plan_set_host_mode(das_plan* plan, int32_t mode);
int64_t das_plan_h2d_explicit_bytes(const das_plan* plan);   /* bytes explicitly copied by the last run_host */
int64_t das_plan_d2h_bytes(const das_plan* plan);
/* after a host_mode-2 run: stats[0] = distinct feature rows the row cache fetched over PCIe, stats[1] = candidates above
 * score_thr (one F(p) row each).  Synchronises with the device; for measurement, not for the hot path. */
int das_plan_row_cache_stats(das_plan* plan, int32_t stats[2]);

/* after a tensor-core-mode run: stats[0] = distinct (cell, joint) feature rows the gathered GEMM multiplied, stats[1] =
 * candidates above score_thr, stats[2] = rows without the de-duplication (valid items * 32).  Synchronises with the
 * device; for measurement (bench.py's roofline), not for the hot path. */
int das_plan_refine_stats(das_plan* plan, int64_t stats[3]);

/* ---- diagnostics ------------------------------------------------------------------------------- */
/* Self-test of the tcgen05/TMEM building blocks: D[128,N] = A[128,K] * B[N,K]^T (row-major fp32 device
 * buffers; N in {16,32}, K a multiple of 32). split=0: one TF32 pass; split=1: 3xTF32 (fp32-level accuracy). */
int das_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t split, void* stream);
/* Which phase 1-2 kernel das_refine_heads launches on device-resident maps: 0 = automatic (warp-per-item below 24*148 items,
 * batched above), 1 = warp-per-item, 2 = batched.  Process-wide; for parity tests and A/B timing.  Plans capture the choice
 * into their graph: set it before the plan's first run. */
int das_debug_force_heads_kernel(int32_t mode);
/* tcgen05.mma issue/throughput micro-benchmark: out_cycles[0] = issue only, [1] = issue + completion (device int64[2]) */
int das_tc_mma_bench(int32_t N, int32_t iters, long long* out_cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAS_DECODE_H_ */
