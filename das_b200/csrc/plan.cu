// Plan API: the whole DASHead.get_poses call (reference das_head.py:653-688) as one CUDA-graph
// replay -- score scan + top-k, (dense layers 1..L-1,) sparse last-layer refinement + assembly,
// OKS-NMS + back-projection -- plus the host-buffer entry used for end-to-end measurements.
#include <algorithm>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "das_common.cuh"

namespace das {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace das

struct das_plan {
    das_decode_cfg cfg{};
    das_levels shape{};        // H, W, stride, batch (pointers unused)
    das_levels bound{};        // last bound inputs (device pointers)
    das_levels* d_levels = nullptr;
    int B = 0, CT = 0, P = 0, hw_sum = 0;
    int device = 0;            // the plan's buffers live here; run/bind must be called with this device current
    das_buffers buf{};
    uint32_t* scratch = nullptr;
    int32_t* work_counter = nullptr;
    float* d_scale_xy = nullptr;
    double* d_cam = nullptr;
    float* wpack[DAS_MAX_LAYERS] = {};
    // tensor-core refinement (C = 256, nh = 4)
    int refine_mode = 0;              // 0 SIMT, 1 tcgen05 3xTF32, 2 tcgen05 single TF32
    int pdl_mode = -1;                // programmatic dependent launch along the kernel chain: 0 off, 1 on, -1 auto (on for small decodes)
    unsigned char* tc_panels = nullptr;
    das_refine_scratch rs{};          // distinct-row lists, row / item records, valid list, counters (= work_counter)
    unsigned char* dense_panels[DAS_MAX_LAYERS] = {};   // tensor-core panels of the dense layers (0..L-2)
    // dense layers (num_layers > 1): ping-pong joint-major [B][J][HW][4] maps per level + projection scratch [B][J][HW][16]
    float* uvd_map[2][DAS_MAX_LEVELS] = {};
    float* proj = nullptr;
    const float** d_prev_ptrs = nullptr;
    // on-demand sampling of layer L-2 (tensor-core path): that layer's projection planes must outlive the other levels'
    // dense layers, so every level has its own (one level: the shared scratch itself)
    float* proj_last[DAS_MAX_LEVELS] = {};
    const float** d_plane_ptrs = nullptr;
    int on_demand = 1;                // das_plan_set_on_demand_sampling
    // host-entry staging
    das_levels staging{};
    bool staging_ready = false;
    int host_mode = 0;                // das_plan_run_host: 0 = bulk H2D of every map, 1 = sparse maps read in place (zero copy),
                                      // 2 = zero copy + device row cache in front of the tensor-core sampling phase
    bool rc_active = false;           // the row-cache pass is part of the enqueued / captured work
    bool in_host_call = false;        // das_plan_bind is being called by das_plan_run_host
    das_row_cache rc{};               // allocated on first use
    int64_t h2d_explicit = 0;         // bytes das_plan_run_host copies explicitly per call in the current host_mode
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    unsigned char* out_block = nullptr;   // all out_* buffers live in this one allocation (one D2H / one all-gather)
    unsigned char* own_block = nullptr;   // the plan's own allocation (out_block may point at a caller-owned block instead)
    size_t out_block_bytes = 0;           // packed outputs; a 256-byte trailer with the sequence word follows
    size_t out_off[7] = {};               // byte offsets of count | score | slot | pose | center | cam | world
    das_peer_blocks peers{};              // fused result all-gather (das_plan_set_peer_blocks)
    // graph
    cudaGraphExec_t exec = nullptr;
    cudaGraphExec_t exec_prof = nullptr;   // same graph with event-record nodes between the stages
    cudaEvent_t ev[DAS_NUM_STAGES + 1] = {};
    cudaStream_t cap_stream = nullptr;   // capture happens here (the caller's stream may be the legacy default stream)
    // result all-gather (peers.n > 0): the publish kernel runs on the plan's own side stream behind the decode, so the
    // caller's stream can start its next decode while the NVLink stores and the system-scope fences are in flight
    cudaStream_t pub_stream = nullptr;
    cudaEvent_t ev_decoded = nullptr, ev_published = nullptr;
    bool pub_pending = false;
    int64_t launches = 0;
    int launches_per_run = 0;
};

extern "C" const char* das_version(void) { return "das-b200 0.1 (sm_100a)"; }
extern "C" const char* das_last_error(void) { return das::g_err; }
extern "C" void das_abi_struct_sizes(int32_t out[6]) {
    out[0] = sizeof(das_levels); out[1] = sizeof(das_decode_cfg); out[2] = sizeof(das_buffers); out[3] = sizeof(das_row_cache);
    out[4] = sizeof(das_refine_scratch); out[5] = sizeof(das_peer_blocks);
}

extern "C" int32_t das_level_slots(int32_t H, int32_t W, int32_t nms_pre) { return das::level_slots(H * W, nms_pre); }

extern "C" int32_t das_candidate_slots(const das_levels* lv, int32_t nms_pre) {
    if (!lv) return 0;
    int t = 0;
    for (int l = 0; l < lv->n_levels; ++l) t += das::level_slots(lv->lv[l].H * lv->lv[l].W, nms_pre);
    return t;
}

static void point_outputs(das_plan* p, unsigned char* q) {
    p->out_block = q;
    p->buf.out_count = reinterpret_cast<int32_t*>(q + p->out_off[0]);
    p->buf.out_score = reinterpret_cast<float*>(q + p->out_off[1]);
    p->buf.out_slot = reinterpret_cast<int32_t*>(q + p->out_off[2]);
    p->buf.out_pose = reinterpret_cast<float*>(q + p->out_off[3]);
    p->buf.out_center = reinterpret_cast<float*>(q + p->out_off[4]);
    p->buf.out_cam = reinterpret_cast<double*>(q + p->out_off[5]);
    p->buf.out_world = reinterpret_cast<double*>(q + p->out_off[6]);
}

template <typename T>
static int dev_alloc(T** p, size_t n) {
    DAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)));
    return DAS_OK;
}

extern "C" int das_plan_create(const das_decode_cfg* cfg, const das_levels* shape, das_plan** out) {
    using namespace das;
    DAS_REQUIRE(cfg && shape && out, DAS_ERR_ARG, "das_plan_create: null pointer");
    DAS_REQUIRE(shape->n_levels >= 1 && shape->n_levels <= DAS_MAX_LEVELS, DAS_ERR_ARG, "n_levels=%d", shape->n_levels);
    DAS_REQUIRE(shape->batch >= 1, DAS_ERR_ARG, "batch=%d", shape->batch);
    DAS_REQUIRE(cfg->num_joints >= 1 && cfg->num_joints <= DAS_MAX_JOINTS, DAS_ERR_CAPACITY, "num_joints=%d (max %d)",
                cfg->num_joints, DAS_MAX_JOINTS);
    DAS_REQUIRE(cfg->root_idx >= 0 && cfg->root_idx < cfg->num_joints, DAS_ERR_ARG, "root_idx=%d", cfg->root_idx);
    DAS_REQUIRE(cfg->nms_pre <= DAS_MAX_NMS_PRE, DAS_ERR_CAPACITY, "nms_pre=%d (max %d)", cfg->nms_pre, DAS_MAX_NMS_PRE);
    if (cfg->refine) {
        DAS_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= DAS_MAX_LAYERS, DAS_ERR_CAPACITY, "num_layers=%d", cfg->num_layers);
        DAS_REQUIRE(cfg->num_heads == 4, DAS_ERR_UNSUPPORTED, "num_heads=%d: only 4 is built", cfg->num_heads);
        DAS_REQUIRE(cfg->feat_channels == 128 || cfg->feat_channels == 256 || cfg->feat_channels == 512, DAS_ERR_UNSUPPORTED,
                    "feat_channels=%d: only 128/256/512 are built", cfg->feat_channels);
    }
    for (int l = 0; l < shape->n_levels; ++l)
        DAS_REQUIRE(shape->lv[l].H > 0 && shape->lv[l].W > 0 && shape->lv[l].stride > 0, DAS_ERR_ARG, "level %d: bad shape", l);
    int device = 0;
    DAS_CUDA_CHECK(cudaGetDevice(&device));
    das_plan* p = new (std::nothrow) das_plan();
    DAS_REQUIRE(p, DAS_ERR_ARG, "out of host memory");
    p->cfg = *cfg;
    p->shape = *shape;
    p->bound = *shape;
    p->B = shape->batch;
    p->device = device;
    for (int l = 0; l < shape->n_levels; ++l) p->hw_sum += shape->lv[l].H * shape->lv[l].W;
    p->CT = das_candidate_slots(shape, cfg->nms_pre);
    p->P = das_output_slots(p->CT, cfg->nms_post);
    if (p->CT > 8192) {
        set_error("candidate slots per image = %d exceed capacity 8192 (lower nms_pre)", p->CT);
        delete p;
        return DAS_ERR_CAPACITY;
    }
    const size_t B = p->B, CT = p->CT, P = p->P, J = cfg->num_joints;
    if (static_cast<long long>(p->B) * p->CT * 32 >= (1ll << 31)) {
        set_error("batch * candidate slots = %lld is beyond the row-record capacity", static_cast<long long>(p->B) * p->CT);
        delete p;
        return DAS_ERR_CAPACITY;
    }
    int s = DAS_OK;
    auto A = [&](int r) { if (s == DAS_OK) s = r; };
    A(dev_alloc(&p->d_levels, 1));
    A(dev_alloc(&p->buf.cand_score, B * CT));
    A(dev_alloc(&p->buf.cand_index, B * CT));
    A(dev_alloc(&p->buf.cand_pose, B * CT * J * 3));
    A(dev_alloc(&p->buf.cand_center, B * CT * 3));
    {
        // carve every output out of one block: [count | score | slot | pose | center | cam | world], 256 B aligned
        auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
        const size_t o_count = 0;
        const size_t o_score = up(o_count + B * 4);
        const size_t o_slot = up(o_score + B * P * 4);
        const size_t o_pose = up(o_slot + B * P * 4);
        const size_t o_center = up(o_pose + B * P * J * 3 * 4);
        const size_t o_cam = up(o_center + B * P * 3 * 4);
        const size_t o_world = up(o_cam + B * P * J * 3 * 8);
        p->out_block_bytes = up(o_world + B * P * J * 3 * 8);
        const size_t offs[7] = {o_count, o_score, o_slot, o_pose, o_center, o_cam, o_world};
        std::memcpy(p->out_off, offs, sizeof(offs));
        A(dev_alloc(&p->own_block, p->out_block_bytes + 256));
        if (s == DAS_OK) {
            if (cudaMemset(p->own_block + p->out_block_bytes, 0, 256) != cudaSuccess) s = DAS_ERR_CUDA;
            point_outputs(p, p->own_block);
        }
    }
    A(dev_alloc(&p->scratch, B * static_cast<size_t>(p->hw_sum)));
    A(dev_alloc(&p->work_counter, 4 + DAS_MAX_JOINTS));
    A(dev_alloc(&p->d_scale_xy, B * 2));
    A(dev_alloc(&p->d_cam, B * DAS_CAM_DOUBLES));
    if (cfg->refine) {
        for (int k = 0; k < cfg->num_layers; ++k) A(dev_alloc(&p->wpack[k], static_cast<size_t>(das_packed_weight_floats(cfg))));
        if (cfg->feat_channels == 256 && cfg->num_heads == 4) {
            // tensor-core refinement (tcgen05 3xTF32, A operand through TMEM) is the default where it is built;
            // das_plan_set_refine_mode(plan, 0) selects the fp32 SIMT kernel
            p->refine_mode = 1;
            A(dev_alloc(&p->tc_panels, static_cast<size_t>(das_tc_panel_bytes(cfg))));
            p->rs.row_cap = static_cast<int32_t>(B * CT * 32);
            A(dev_alloc(&p->rs.unique_rows, J * B * CT * 32 * 8));
            A(dev_alloc(&p->rs.unique_out, J * B * CT * 32 * 8));
            A(dev_alloc(&p->rs.row_records, B * CT * J * 32 * 4));
            A(dev_alloc(&p->rs.item_records, B * CT * J * 8));
            A(dev_alloc(&p->rs.valid_list, B * CT));
            p->rs.counters = p->work_counter;
        }
        if (cfg->num_layers > 1) {
            size_t max_hw = 0;
            for (int l = 0; l < shape->n_levels; ++l) {
                const size_t hw = static_cast<size_t>(shape->lv[l].H) * shape->lv[l].W;
                max_hw = std::max(max_hw, hw);
                A(dev_alloc(&p->uvd_map[0][l], B * hw * 4 * J));          // joint-major [B][J][HW][4]
                if (cfg->num_layers > 2) A(dev_alloc(&p->uvd_map[1][l], B * hw * 4 * J));
            }
            A(dev_alloc(&p->proj, B * max_hw * (2 * cfg->num_heads + 8) * J));
            if (cfg->feat_channels == 256 && cfg->num_heads == 4)
                for (int k = 0; k < cfg->num_layers - 1; ++k) A(dev_alloc(&p->dense_panels[k], static_cast<size_t>(das_dense_panel_bytes(cfg))));
            A(dev_alloc(&p->d_prev_ptrs, DAS_MAX_LEVELS));
            if (cfg->feat_channels == 256 && cfg->num_heads == 4) {
                A(dev_alloc(&p->d_plane_ptrs, DAS_MAX_LEVELS));
                for (int l = 0; l < shape->n_levels && shape->n_levels > 1; ++l)
                    A(dev_alloc(&p->proj_last[l], B * static_cast<size_t>(shape->lv[l].H) * shape->lv[l].W * (2 * cfg->num_heads + 8) * J));
            }
        }
    }
    if (s != DAS_OK) { das_plan_destroy(p); return s; }
    // identity metas until das_plan_set_metas is called
    {
        float* sxy = new float[B * 2];
        double* cam = new double[B * DAS_CAM_DOUBLES];
        for (size_t b = 0; b < B; ++b) {
            sxy[2 * b] = sxy[2 * b + 1] = 1.f;
            const double ident[DAS_CAM_DOUBLES] = {1, 0, 0, 0, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
            std::memcpy(cam + b * DAS_CAM_DOUBLES, ident, sizeof(ident));
        }
        cudaError_t e1 = cudaMemcpy(p->d_scale_xy, sxy, B * 2 * sizeof(float), cudaMemcpyHostToDevice);
        cudaError_t e2 = cudaMemcpy(p->d_cam, cam, B * DAS_CAM_DOUBLES * sizeof(double), cudaMemcpyHostToDevice);
        delete[] sxy;
        delete[] cam;
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("das_plan_create: meta upload failed");
            das_plan_destroy(p);
            return DAS_ERR_CUDA;
        }
    }
    p->h2d_bytes = 0;
    for (int l = 0; l < shape->n_levels; ++l) {
        const int64_t hw = static_cast<int64_t>(shape->lv[l].H) * shape->lv[l].W;
        p->h2d_bytes += static_cast<int64_t>(B) * hw * 4 * (2 + 3 + 6 * J);
        if (cfg->refine) p->h2d_bytes += static_cast<int64_t>(B) * hw * 4 * cfg->feat_channels * cfg->num_layers;
    }
    p->h2d_bytes += static_cast<int64_t>(B) * (2 * 4 + DAS_CAM_DOUBLES * 8);
    p->d2h_bytes = static_cast<int64_t>(B) * 4 + static_cast<int64_t>(B) * P * (4 + 4 + 12 + J * 3 * (4 + 8 + 8));
    *out = p;
    return DAS_OK;
}

extern "C" void das_plan_destroy(das_plan* p) {
    if (!p) return;
    if (p->exec) cudaGraphExecDestroy(p->exec);
    if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
    if (p->pub_stream) cudaStreamDestroy(p->pub_stream);
    if (p->ev_decoded) cudaEventDestroy(p->ev_decoded);
    if (p->ev_published) cudaEventDestroy(p->ev_published);
    if (p->exec_prof) cudaGraphExecDestroy(p->exec_prof);
    for (cudaEvent_t e : p->ev) if (e) cudaEventDestroy(e);
    void* ptrs[] = {p->d_levels, p->buf.cand_score, p->buf.cand_index, p->buf.cand_pose, p->buf.cand_center,
                    p->own_block, p->scratch, p->work_counter, p->d_scale_xy, p->d_cam,
                    p->proj, p->d_prev_ptrs, p->d_plane_ptrs, p->tc_panels, p->rs.unique_rows, p->rs.unique_out, p->rs.row_records,
                    p->rs.item_records, p->rs.valid_list};
    for (void* q : ptrs) if (q) cudaFree(q);
    if (p->rc.table) cudaFree(p->rc.table);
    if (p->rc.rows) cudaFree(p->rc.rows);
    if (p->rc.cand_rows) cudaFree(p->rc.cand_rows);
    for (int k = 0; k < DAS_MAX_LAYERS; ++k) if (p->wpack[k]) cudaFree(p->wpack[k]);
    for (int k = 0; k < DAS_MAX_LAYERS; ++k) if (p->dense_panels[k]) cudaFree(p->dense_panels[k]);
    for (int l = 0; l < DAS_MAX_LEVELS; ++l) if (p->proj_last[l]) cudaFree(p->proj_last[l]);
    for (int i = 0; i < 2; ++i)
        for (int l = 0; l < DAS_MAX_LEVELS; ++l) if (p->uvd_map[i][l]) cudaFree(p->uvd_map[i][l]);
    if (p->staging_ready) {
        for (int l = 0; l < p->shape.n_levels; ++l) {
            cudaFree(const_cast<float*>(p->staging.lv[l].cls));
            cudaFree(const_cast<float*>(p->staging.lv[l].ctr));
            cudaFree(const_cast<float*>(p->staging.lv[l].pose));
            for (int k = 0; k < DAS_MAX_LAYERS; ++k)
                if (p->staging.lv[l].feats[k]) cudaFree(const_cast<float*>(p->staging.lv[l].feats[k]));
        }
    }
    delete p;
}

extern "C" int das_plan_set_weights(das_plan* p, int32_t layer, const float* so_w, const float* so_b,
                                    const float* sc_w, const float* sc_b, const float* uw_w, const float* uw_b,
                                    const float* uv_w, const float* uv_b, void* stream) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    DAS_REQUIRE(p->cfg.refine, DAS_ERR_ARG, "plan was created with refine=0");
    DAS_REQUIRE(layer >= 0 && layer < p->cfg.num_layers, DAS_ERR_ARG, "layer=%d of %d", layer, p->cfg.num_layers);
    DAS_TRY(das_pack_weights(&p->cfg, so_w, so_b, sc_w, sc_b, uw_w, uw_b, uv_w, uv_b, p->wpack[layer], stream));
    if (layer == p->cfg.num_layers - 1 && p->tc_panels) DAS_TRY(das_pack_tc_panels(&p->cfg, p->wpack[layer], p->tc_panels, stream));
    if (layer < p->cfg.num_layers - 1 && p->dense_panels[layer]) DAS_TRY(das_pack_dense_panels(&p->cfg, p->wpack[layer], p->dense_panels[layer], stream));
    return DAS_OK;
}

static int set_row_cache(das_plan* p, bool on);

static void drop_graphs(das_plan* p);

extern "C" int das_plan_bind(das_plan* p, const das_levels* levels, void* stream) {
    using namespace das;
    DAS_REQUIRE(p && levels, DAS_ERR_ARG, "das_plan_bind: null pointer");
    if (!p->in_host_call) DAS_TRY(set_row_cache(p, false));   // device-resident inputs: L2 already gives the row re-use
    DAS_REQUIRE(levels->n_levels == p->shape.n_levels && levels->batch == p->shape.batch, DAS_ERR_ARG,
                "bind: n_levels/batch (%d/%d) differ from the plan (%d/%d)", levels->n_levels, levels->batch,
                p->shape.n_levels, p->shape.batch);
    DAS_REQUIRE(levels->in_dtype == DAS_DTYPE_F32 || levels->in_dtype == DAS_DTYPE_F16 || levels->in_dtype == DAS_DTYPE_BF16, DAS_ERR_ARG,
                "bind: in_dtype=%d (DAS_DTYPE_F32 / F16 / BF16)", levels->in_dtype);
    for (int l = 0; l < levels->n_levels; ++l) {
        const das_level_desc& d = levels->lv[l];
        DAS_REQUIRE(d.H == p->shape.lv[l].H && d.W == p->shape.lv[l].W && d.stride == p->shape.lv[l].stride, DAS_ERR_ARG,
                    "bind: level %d shape differs from the plan", l);
        DAS_REQUIRE((reinterpret_cast<uintptr_t>(d.cls) | reinterpret_cast<uintptr_t>(d.ctr) | reinterpret_cast<uintptr_t>(d.pose)) %
                    das::dtype_bytes(levels->in_dtype) == 0, DAS_ERR_ARG, "bind: level %d map is not aligned to its element type", l);
        DAS_REQUIRE(d.cls && d.ctr && d.pose, DAS_ERR_ARG, "bind: level %d has a null map", l);
        if (p->cfg.refine)
            for (int k = 0; k < p->cfg.num_layers; ++k) {
                DAS_REQUIRE(d.feats[k], DAS_ERR_ARG, "bind: level %d layer %d feature map is null", l, k);
                DAS_REQUIRE((reinterpret_cast<uintptr_t>(d.feats[k]) & 15) == 0, DAS_ERR_ARG, "feature maps must be 16-byte aligned");
            }
    }
    if (levels->in_dtype != p->bound.in_dtype) drop_graphs(p);   // score_topk is instantiated per element type: re-capture
    p->bound = *levels;
    // pageable host -> device copy of 0.5 KB: stream-ordered, returns after staging
    DAS_CUDA_CHECK(cudaMemcpyAsync(p->d_levels, &p->bound, sizeof(das_levels), cudaMemcpyHostToDevice,
                                   static_cast<cudaStream_t>(stream)));
    return DAS_OK;
}

extern "C" int das_plan_set_metas(das_plan* p, const float* scale_xy, const double* cam, void* stream) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (scale_xy) DAS_CUDA_CHECK(cudaMemcpyAsync(p->d_scale_xy, scale_xy, sizeof(float) * 2 * p->B, cudaMemcpyHostToDevice, st));
    if (cam) DAS_CUDA_CHECK(cudaMemcpyAsync(p->d_cam, cam, sizeof(double) * DAS_CAM_DOUBLES * p->B, cudaMemcpyHostToDevice, st));
    return DAS_OK;
}

namespace das {
ChainCtx& chain_ctx() {
    static thread_local ChainCtx ctx;
    return ctx;
}
}  // namespace das

// RAII: the stage launchers called inside see the plan's chain context, everybody else the stand-alone default
struct ChainScope {
    das::ChainCtx saved;
    explicit ChainScope(const das::ChainCtx& c) : saved(das::chain_ctx()) { das::chain_ctx() = c; }
    ~ChainScope() { das::chain_ctx() = saved; }
};

static bool peer_inline() {
    static const bool v = std::getenv("DAS_PEER_INLINE") && std::getenv("DAS_PEER_INLINE")[0] == '1';
    return v;
}
static bool peer_fused() {
    static const bool v = std::getenv("DAS_PEER_FUSED") && std::getenv("DAS_PEER_FUSED")[0] == '1';
    return v;
}

// Layer L-2's progressive sampling evaluated on demand by the sparse last layer (tensor-core path, default kernels)?
static bool on_demand_sampling(const das_plan* p) {
    static const bool env_off = (std::getenv("DAS_ON_DEMAND") && std::getenv("DAS_ON_DEMAND")[0] == '0') ||
                                (std::getenv("DAS_HEADS_SPLIT") && std::getenv("DAS_HEADS_SPLIT")[0] == '0') ||
                                (std::getenv("DAS_HEADS_NB") && std::atoi(std::getenv("DAS_HEADS_NB")) != 4);
    const das_decode_cfg& c = p->cfg;
    return !env_off && p->on_demand && c.refine && c.num_layers > 1 && p->refine_mode != 0 && p->d_plane_ptrs &&
           p->dense_panels[c.num_layers - 2];
}

static int enqueue(das_plan* p, cudaStream_t st, int* n_launch, bool events) {
    const das_decode_cfg& c = p->cfg;
    int n = 0;
    // Programmatic dependent launch along the chain (DAS_PDL=0 switches it off).  Not in the stage-timing variant: its
    // event-record nodes sit between the kernels, and a timed stage should not overlap its neighbours anyway.
    static const bool pdl_env = !(std::getenv("DAS_PDL") && std::getenv("DAS_PDL")[0] == '0');
    // pdl_mode: early-launched dependents wait on the SMs they will run on.  On one stream that hides every launch gap and
    // ramp (B=64: 105 -> 81 us per decode, B=1: 37 us); with several independent decodes in flight the waiting CTAs take SM
    // resources from the other streams' kernels (-4 % throughput), so "auto" enables it for decodes that cannot fill the GPU.
    const long long items = static_cast<long long>(p->B) * p->CT * c.num_joints;
    const bool pdl_want = p->pdl_mode == 1 || (p->pdl_mode < 0 && items <= 24LL * das::kSMs);
    const bool pdl_on = pdl_env && !events && pdl_want;
    das::ChainCtx cx;
    cx.pdl = false;              // the first kernel of the chain has no kernel of THIS decode in front of it
    const bool tc_chain = c.refine && p->refine_mode != 0 && !p->rc_active;
    if (pdl_on && tc_chain) { cx.zero_counters = p->work_counter; cx.counters_cleared = true; }
    ChainScope scope(cx);
    auto mark = [&](int i) -> int {
        // external: becomes a real event-record node when captured, so cudaEventElapsedTime works after a replay
        if (events) DAS_CUDA_CHECK(cudaEventRecordWithFlags(p->ev[i], st, cudaEventRecordExternal));
        return DAS_OK;
    };
    DAS_TRY(mark(0));
    DAS_TRY(das_score_topk(p->d_levels, &p->bound, c.nms_pre, c.peak_kernel, p->buf.cand_score, p->buf.cand_index,
                           p->CT, p->scratch, st));
    ++n;
    das::chain_ctx().pdl = pdl_on;
    DAS_TRY(mark(1));
    const float* const* prev = nullptr;
    const bool lazy = on_demand_sampling(p);
    if (c.refine && c.num_layers > 1) {
        for (int l = 0; l < p->bound.n_levels; ++l) {
            const float* in = nullptr;
            for (int k = 0; k < c.num_layers - 1; ++k) {
                if (lazy && k == c.num_layers - 2) {
                    // projection only: the last layer samples these planes at the few cells it looks at (das_refine_heads)
                    DAS_TRY(das_dense_project_tc(p->d_levels, &p->bound, l, k, &c, p->wpack[k], p->dense_panels[k], in,
                                                 p->bound.n_levels > 1 ? p->proj_last[l] : p->proj, st));
                    n += 1;
                    break;
                }
                float* outm = p->uvd_map[k & 1][l];
                DAS_TRY(das_refine_dense_layer(p->d_levels, &p->bound, l, k, &c, p->wpack[k],
                                               p->refine_mode != 0 ? p->dense_panels[k] : nullptr, in, outm, p->proj, st));
                n += 2;
                in = outm;
            }
        }
        prev = p->d_prev_ptrs;   // uploaded once in das_plan_run: uvd_map[(L-2)&1][level]
    }
    DAS_TRY(mark(2));
    if (c.refine && p->refine_mode != 0) {
        const float* w = p->wpack[c.num_layers - 1];
        if (p->rc_active) {         // host zero-copy mode: every distinct feature row crosses PCIe once
            DAS_TRY(das_row_cache_clear(&p->rc, st));
            DAS_TRY(das_refine_cand_rows(p->d_levels, &p->bound, &c, p->buf.cand_score, p->buf.cand_index, p->CT, &p->rc, st));
            ++n;
        }
        das_refine_scratch rs_heads = p->rs;
        rs_heads.prev_planes = lazy ? p->d_plane_ptrs : nullptr;
        DAS_TRY(das_refine_heads(p->d_levels, &p->bound, &c, w, prev, p->d_scale_xy, p->buf.cand_score, p->buf.cand_index, p->CT,
                                 &rs_heads, p->buf.cand_center, p->rc_active ? &p->rc : nullptr, st));
        DAS_TRY(mark(3));
        if (p->rc_active) {
            DAS_TRY(das_refine_row_cache(&c, &p->rs, &p->rc, st));
            ++n;
        }
        DAS_TRY(das_refine_tc(p->d_levels, &p->bound, &c, w, p->tc_panels, p->CT, &p->rs, p->refine_mode == 1 ? 1 : 0, st));
        DAS_TRY(das_refine_finish(&p->bound, &c, p->CT, &p->rs, p->buf.cand_pose, st));
        ++n;
        n += 2;
    } else {
        DAS_TRY(das_gather_refine_assemble(p->d_levels, &p->bound, &c, c.refine ? p->wpack[c.num_layers - 1] : nullptr, prev,
                                           p->d_scale_xy, p->buf.cand_score, p->buf.cand_index, p->CT, p->buf.cand_pose,
                                           p->buf.cand_center, p->work_counter, st));
        ++n;
        DAS_TRY(mark(3));
    }
    DAS_TRY(mark(4));
    // result all-gather: a small publish kernel behind the NMS kernel (default), or the stores fused into the NMS kernel's
    // own CTAs (DAS_PEER_FUSED=1; measured slower, see das_peer_publish)
    const bool fused_peers = peer_fused();
    DAS_TRY(das_nms_backproject_peers(&c, p->B, p->CT, p->buf.cand_score, p->buf.cand_pose, p->buf.cand_center, p->d_cam,
                                      p->buf, (p->peers.n > 0 && fused_peers) ? &p->peers : nullptr, st));
    ++n;
    // DAS_PEER_INLINE=1: the publish kernel as the last node of the chain (the next decode on this stream waits for its NVLink
    // round trips: +3.9 us per step at N = 2); default: on the plan's side stream, see das_plan_run
    if (p->peers.n > 0 && !fused_peers && peer_inline()) {
        DAS_TRY(das_peer_publish(&p->peers, p->out_block, static_cast<int64_t>((p->out_block_bytes + 15) & ~static_cast<size_t>(15)), st));
        ++n;
    }
    DAS_TRY(mark(5));
    n += das::chain_ctx().extra_launches;
    *n_launch = n;
    return DAS_OK;
}

static int capture(das_plan* p, cudaGraphExec_t* exec, bool events) {
    cudaGraph_t g = nullptr;
    int n = 0;
    if (!p->cap_stream) DAS_CUDA_CHECK(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
    DAS_CUDA_CHECK(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal));
    int s = enqueue(p, p->cap_stream, &n, events);
    cudaError_t e = cudaStreamEndCapture(p->cap_stream, &g);
    if (s != DAS_OK) { if (g) cudaGraphDestroy(g); return s; }
    DAS_CUDA_CHECK(e);
    e = cudaGraphInstantiate(exec, g, 0);
    cudaGraphDestroy(g);
    DAS_CUDA_CHECK(e);
    return DAS_OK;
}

// mode: 0 = eager launches, 1 = CUDA graph replay, 2 = graph replay with stage-boundary events
extern "C" int das_plan_run(das_plan* p, void* stream, int32_t mode) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    DAS_REQUIRE(mode >= 0 && mode <= 2, DAS_ERR_ARG, "mode=%d", mode);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int cur_device = -1;
    DAS_CUDA_CHECK(cudaGetDevice(&cur_device));
    DAS_REQUIRE(cur_device == p->device, DAS_ERR_ARG, "plan was created on device %d but device %d is current", p->device, cur_device);
    if (p->launches == 0) {
        if (p->cfg.refine && p->cfg.num_layers > 1) {
            const float* host_prev[DAS_MAX_LEVELS] = {};
            for (int l = 0; l < p->bound.n_levels; ++l) host_prev[l] = p->uvd_map[(p->cfg.num_layers - 2) & 1][l];
            DAS_CUDA_CHECK(cudaMemcpy(p->d_prev_ptrs, host_prev, sizeof(host_prev), cudaMemcpyHostToDevice));
            if (p->d_plane_ptrs) {
                const float* host_planes[DAS_MAX_LEVELS] = {};
                for (int l = 0; l < p->bound.n_levels; ++l) host_planes[l] = p->bound.n_levels > 1 ? p->proj_last[l] : p->proj;
                DAS_CUDA_CHECK(cudaMemcpy(p->d_plane_ptrs, host_planes, sizeof(host_planes), cudaMemcpyHostToDevice));
            }
        }
        for (cudaEvent_t& e : p->ev) DAS_CUDA_CHECK(cudaEventCreate(&e));
    }
    const bool side_publish = p->peers.n > 0 && !peer_fused() && !peer_inline();
    if (side_publish && p->pub_pending) {
        // the previous publish of THIS plan has read the output block this decode is about to overwrite (long done in practice)
        DAS_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_published, 0));
    }
    struct Publish {          // enqueued behind whatever decode this call launches, on every return path
        das_plan* p; cudaStream_t st; bool on;
        int go() {
            if (!on) return DAS_OK;
            if (!p->pub_stream) {
                DAS_CUDA_CHECK(cudaStreamCreateWithFlags(&p->pub_stream, cudaStreamNonBlocking));
                DAS_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_decoded, cudaEventDisableTiming));
                DAS_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_published, cudaEventDisableTiming));
            }
            DAS_CUDA_CHECK(cudaEventRecord(p->ev_decoded, st));
            DAS_CUDA_CHECK(cudaStreamWaitEvent(p->pub_stream, p->ev_decoded, 0));
            DAS_TRY(das_peer_publish(&p->peers, p->out_block, static_cast<int64_t>((p->out_block_bytes + 15) & ~static_cast<size_t>(15)), p->pub_stream));
            DAS_CUDA_CHECK(cudaEventRecord(p->ev_published, p->pub_stream));
            p->pub_pending = true;
            p->launches += 1;
            return DAS_OK;
        }
    } publish{p, st, side_publish};
    if (mode == 0 || p->launches == 0) {
        // the very first run is always eager (module loading and cudaFuncSetAttribute must not land inside a capture);
        // the graph is captured by the next call, so no call runs the decode twice
        int n = 0;
        DAS_TRY(enqueue(p, st, &n, false));
        p->launches_per_run = n;
        p->launches += n;
        if (mode != 2) return publish.go();
        DAS_CUDA_CHECK(cudaStreamSynchronize(st));      // profiling replay requested on a fresh plan: capture right away
    }
    cudaGraphExec_t* ex = (mode == 2) ? &p->exec_prof : &p->exec;
    if (!*ex) DAS_TRY(capture(p, ex, mode == 2));
    DAS_CUDA_CHECK(cudaGraphLaunch(*ex, st));
    p->launches += p->launches_per_run;
    return publish.go();
}

// Makes `stream` wait until the plan's last result publication to its peers (das_plan_set_peer_blocks) has been issued and
// completed on the plan's side stream; a no-op without peers.  Call it where the caller needs "every peer has my results".
extern "C" int das_plan_publish_wait(das_plan* p, void* stream) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    if (p->pub_pending) DAS_CUDA_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), p->ev_published, 0));
    return DAS_OK;
}

// Milliseconds of each stage of the last mode-2 replay: {score_topk, dense layers, refine phases 1-2 (0 in SIMT mode),
// refine + assemble, nms+backproject}.
// The caller must have synchronised the stream.
extern "C" int das_plan_stage_ms(das_plan* p, float* ms) {
    using namespace das;
    DAS_REQUIRE(p && ms, DAS_ERR_ARG, "das_plan_stage_ms: null pointer");
    DAS_REQUIRE(p->exec_prof, DAS_ERR_ARG, "das_plan_stage_ms: no mode-2 run yet");
    for (int i = 0; i < DAS_NUM_STAGES; ++i) DAS_CUDA_CHECK(cudaEventElapsedTime(&ms[i], p->ev[i], p->ev[i + 1]));
    return DAS_OK;
}

extern "C" int das_plan_output_block(const das_plan* p, void** ptr, int64_t* bytes) {
    using namespace das;
    DAS_REQUIRE(p && ptr && bytes, DAS_ERR_ARG, "das_plan_output_block: null pointer");
    *ptr = p->out_block;
    *bytes = static_cast<int64_t>(p->out_block_bytes);
    return DAS_OK;
}

static void drop_graphs(das_plan* p) {
    if (p->exec) { cudaGraphExecDestroy(p->exec); p->exec = nullptr; }
    if (p->exec_prof) { cudaGraphExecDestroy(p->exec_prof); p->exec_prof = nullptr; }
}

extern "C" int das_plan_set_output_block(das_plan* p, void* block, int64_t bytes) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    if (!block) {                                  // back to the plan's own allocation
        point_outputs(p, p->own_block);
    } else {
        DAS_REQUIRE((reinterpret_cast<uintptr_t>(block) & 255) == 0, DAS_ERR_ARG, "output block must be 256-byte aligned");
        DAS_REQUIRE(bytes >= static_cast<int64_t>(p->out_block_bytes + 256), DAS_ERR_ARG,
                    "output block of %lld bytes is smaller than the %lld the plan needs", static_cast<long long>(bytes),
                    static_cast<long long>(p->out_block_bytes + 256));
        DAS_CUDA_CHECK(cudaMemset(static_cast<unsigned char*>(block) + p->out_block_bytes, 0, 256));
        point_outputs(p, static_cast<unsigned char*>(block));
    }
    p->peers = das_peer_blocks{};                  // peer deltas were relative to the old block
    drop_graphs(p);
    return DAS_OK;
}

extern "C" int das_plan_set_peer_blocks(das_plan* p, int32_t n_peers, void* const* peer_blocks) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    DAS_REQUIRE(n_peers >= 0 && n_peers <= DAS_MAX_PEERS && (n_peers == 0 || peer_blocks), DAS_ERR_ARG, "n_peers=%d (max %d)",
                n_peers, DAS_MAX_PEERS);
    das_peer_blocks pb{};
    pb.n = n_peers;
    for (int q = 0; q < n_peers; ++q) {
        DAS_REQUIRE(peer_blocks[q] && (reinterpret_cast<uintptr_t>(peer_blocks[q]) & 255) == 0, DAS_ERR_ARG, "peer block %d is null or unaligned", q);
        pb.delta[q] = static_cast<int64_t>(reinterpret_cast<intptr_t>(peer_blocks[q]) - reinterpret_cast<intptr_t>(p->out_block));
    }
    if (n_peers > 0) {
        pb.seq = reinterpret_cast<int32_t*>(p->out_block + p->out_block_bytes);
        pb.ticket = p->work_counter + 2;
        DAS_CUDA_CHECK(cudaMemset(pb.ticket, 0, sizeof(int32_t)));
    }
    p->peers = pb;
    drop_graphs(p);
    return DAS_OK;
}

extern "C" int das_ipc_alloc(int64_t bytes, void** dev_ptr, unsigned char handle[64]) {
    using namespace das;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    DAS_REQUIRE(bytes > 0 && dev_ptr && handle, DAS_ERR_ARG, "das_ipc_alloc: bad argument");
    void* q = nullptr;
    DAS_CUDA_CHECK(cudaMalloc(&q, static_cast<size_t>(bytes)));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, q);
    if (e != cudaSuccess) { cudaFree(q); set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return DAS_ERR_CUDA; }
    DAS_CUDA_CHECK(cudaMemset(q, 0, static_cast<size_t>(bytes)));
    std::memcpy(handle, &h, 64);
    *dev_ptr = q;
    return DAS_OK;
}
extern "C" int das_ipc_open(const unsigned char handle[64], void** dev_ptr) {
    using namespace das;
    DAS_REQUIRE(handle && dev_ptr, DAS_ERR_ARG, "das_ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    DAS_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DAS_OK;
}
extern "C" int das_ipc_close(void* dev_ptr) {
    using namespace das;
    if (dev_ptr) DAS_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
    return DAS_OK;
}
extern "C" int das_ipc_free(void* dev_ptr) {
    using namespace das;
    if (dev_ptr) DAS_CUDA_CHECK(cudaFree(dev_ptr));
    return DAS_OK;
}

extern "C" int das_plan_set_pdl(das_plan* p, int32_t mode) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    DAS_REQUIRE(mode >= -1 && mode <= 1, DAS_ERR_ARG, "das_plan_set_pdl: mode=%d (-1 auto, 0 off, 1 on)", mode);
    if (mode != p->pdl_mode) drop_graphs(p);
    p->pdl_mode = mode;
    return DAS_OK;
}

extern "C" int das_plan_set_on_demand_sampling(das_plan* p, int32_t on) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    DAS_REQUIRE(on == 0 || on == 1, DAS_ERR_ARG, "das_plan_set_on_demand_sampling: on=%d", on);
    if (on != p->on_demand) drop_graphs(p);
    p->on_demand = on;
    return DAS_OK;
}

extern "C" int das_plan_set_refine_mode(das_plan* p, int32_t mode) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    DAS_REQUIRE(mode >= 0 && mode <= 2, DAS_ERR_ARG, "refine mode %d (0 = fp32 SIMT, 1 = tcgen05 3xTF32, 2 = tcgen05 TF32)", mode);
    DAS_REQUIRE(p->launches == 0, DAS_ERR_ARG, "das_plan_set_refine_mode must be called before the first run");
    DAS_REQUIRE(mode == 0 || p->tc_panels, DAS_ERR_UNSUPPORTED, "tensor-core refinement needs refine=1, feat_channels=256, num_heads=4");
    p->refine_mode = mode;
    return DAS_OK;
}

extern "C" int das_plan_buffers(const das_plan* p, das_buffers* out, int32_t* cand_slots, int32_t* out_slots) {
    using namespace das;
    DAS_REQUIRE(p && out, DAS_ERR_ARG, "das_plan_buffers: null pointer");
    *out = p->buf;
    if (cand_slots) *cand_slots = p->CT;
    if (out_slots) *out_slots = p->P;
    return DAS_OK;
}

extern "C" int64_t das_plan_kernel_launches(const das_plan* p) { return p ? p->launches : 0; }
extern "C" int64_t das_plan_h2d_bytes(const das_plan* p) { return p ? p->h2d_bytes : 0; }
// bytes the last das_plan_run_host call moved with explicit copies (host_mode 1 reads the sparse maps in place)
extern "C" int64_t das_plan_h2d_explicit_bytes(const das_plan* p) { return p ? p->h2d_explicit : 0; }
extern "C" int das_plan_set_host_mode(das_plan* p, int32_t mode) {
    using namespace das;
    DAS_REQUIRE(p, DAS_ERR_ARG, "null plan");
    DAS_REQUIRE(mode >= 0 && mode <= 2, DAS_ERR_ARG, "host mode %d", mode);
    p->host_mode = mode;
    return DAS_OK;
}

// Switch the row-cache pass on/off for the following runs (buffers are allocated on first use; a captured graph that
// was built with the other setting is dropped and re-captured by the next das_plan_run).
static int set_row_cache(das_plan* p, bool on) {
    using namespace das;
    on = on && p->cfg.refine && p->refine_mode != 0 && p->tc_panels;
    if (on && !p->rc.table) {
        // distinct rows are bounded both by the row records (32 sampled + 4 target-corner rows per item) and by the
        // number of cells there are; the table holds twice that many entries
        const long long records = std::min<long long>(static_cast<long long>(p->B) * p->CT * p->cfg.num_joints * (32 + 4),
                                                      static_cast<long long>(p->B) * p->hw_sum);
        int bits = 10;
        while (bits < 26 && (1ll << bits) < 2 * records) ++bits;
        long long cap = std::min<long long>(records, 262144);
        if (const char* e = std::getenv("DAS_ROW_CACHE_ROWS")) cap = std::max<long long>(1, std::min<long long>(records, std::atoll(e)));
        unsigned char* t = nullptr;
        DAS_TRY(dev_alloc(&t, static_cast<size_t>(das_row_cache_table_bytes(bits))));
        p->rc.table = t;
        DAS_TRY(dev_alloc(&p->rc.rows, static_cast<size_t>(cap) * p->cfg.feat_channels));
        DAS_TRY(dev_alloc(&p->rc.cand_rows, static_cast<size_t>(p->B) * p->CT * p->cfg.feat_channels));
        p->rc.table_bits = bits;
        p->rc.max_rows = static_cast<int>(cap);
    }
    if (on != p->rc_active) {
        p->rc_active = on;
        drop_graphs(p);
    }
    return DAS_OK;
}
extern "C" int64_t das_plan_d2h_bytes(const das_plan* p) { return p ? p->d2h_bytes : 0; }

// Row-cache occupancy after the last host-mode run: stats[0] = distinct feature rows copied into the device row buffer,
// stats[1] = candidates above score_thr (each fetched its F(p) row once).  Synchronises with the device.
extern "C" int das_plan_row_cache_stats(das_plan* p, int32_t stats[2]) {
    using namespace das;
    DAS_REQUIRE(p && stats, DAS_ERR_ARG, "das_plan_row_cache_stats: null pointer");
    stats[0] = stats[1] = 0;
    if (p->rc.table) {
        const size_t n = static_cast<size_t>(1) << p->rc.table_bits;
        const unsigned char* counter = static_cast<const unsigned char*>(p->rc.table) + n * 12;
        int32_t c = -1;
        DAS_CUDA_CHECK(cudaMemcpy(&c, counter, sizeof(c), cudaMemcpyDeviceToHost));
        stats[0] = std::min<int32_t>(c + 1, p->rc.max_rows);
    }
    int32_t v = 0;
    DAS_CUDA_CHECK(cudaMemcpy(&v, p->work_counter + 1, sizeof(v), cudaMemcpyDeviceToHost));
    stats[1] = v;
    return DAS_OK;
}

// Sparse-refinement counters of the last run (tensor-core mode): stats[0] = DISTINCT (cell, joint) feature rows the
// gathered GEMM multiplied (the bytes das_refine_tc really has to gather: stats[0] * feat_channels * 4), stats[1] =
// candidates above score_thr, stats[2] = row slots it would take without the de-duplication (valid items * 32).
// Synchronises with the device; for measurement, not for the hot path.
extern "C" int das_plan_refine_stats(das_plan* p, int64_t stats[3]) {
    using namespace das;
    DAS_REQUIRE(p && stats, DAS_ERR_ARG, "das_plan_refine_stats: null pointer");
    stats[0] = stats[1] = stats[2] = 0;
    if (!p->work_counter) return DAS_OK;
    int32_t c[4 + DAS_MAX_JOINTS];
    DAS_CUDA_CHECK(cudaDeviceSynchronize());
    DAS_CUDA_CHECK(cudaMemcpy(c, p->work_counter, sizeof(c), cudaMemcpyDeviceToHost));
    for (int j = 0; j < p->cfg.num_joints; ++j) stats[0] += c[4 + j];
    stats[1] = c[1];
    stats[2] = static_cast<int64_t>(c[1]) * p->cfg.num_joints * 32;
    return DAS_OK;
}

extern "C" int das_plan_run_host(das_plan* p, const das_levels* levels, const float* scale_xy, const double* cam,
                                 das_buffers host_out, void* stream) {
    using namespace das;
    DAS_TRY(das_plan_run_host_async(p, levels, scale_xy, cam, host_out, stream));
    DAS_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return DAS_OK;
}

extern "C" int das_plan_run_host_async(das_plan* p, const das_levels* levels, const float* scale_xy, const double* cam,
                                       das_buffers host_out, void* stream) {
    using namespace das;
    DAS_REQUIRE(p && levels, DAS_ERR_ARG, "das_plan_run_host: null pointer");
    DAS_REQUIRE(levels->n_levels == p->shape.n_levels && levels->batch == p->shape.batch, DAS_ERR_ARG,
                "run_host: n_levels/batch differ from the plan");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t B = p->B, J = p->cfg.num_joints, C = p->cfg.feat_channels;
    const int L = p->cfg.refine ? p->cfg.num_layers : 0;
    // host_mode 1: the decode touches every cell of the centre / centerness logits but only ~5 % of the pose and
    // feature maps (K cells and their bilinear corners per image).  Pinned host memory is device-addressable
    // (unified addressing), so those sparse maps are read IN PLACE over PCIe by the gather kernels instead of being
    // copied wholesale; only the logit planes are staged.
    bool zero_copy = p->host_mode >= 1;
    // dense layers (num_layers > 1) read every cell of the pose map and of the feature maps of layers 0..L-2: those
    // are bulk-copied in either mode; only maps that are touched sparsely are read in place
    const bool pose_sparse = L <= 1;
    const float* dev_alias[DAS_MAX_LEVELS][1 + DAS_MAX_LAYERS] = {};
    if (zero_copy) {
        for (int l = 0; l < p->shape.n_levels && zero_copy; ++l) {
            const das_level_desc& h = levels->lv[l];
            const void* srcs[1 + DAS_MAX_LAYERS] = {h.pose};
            for (int k = 0; k < L; ++k) srcs[1 + k] = h.feats[k];
            for (int k = 0; k < 1 + L; ++k) {
                cudaPointerAttributes at{};
                if (!srcs[k] || cudaPointerGetAttributes(&at, srcs[k]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
                    cudaGetLastError();
                    zero_copy = false;      // pageable memory: fall back to staging copies
                    break;
                }
                dev_alias[l][k] = static_cast<const float*>(at.devicePointer);
            }
        }
    }
    if (!p->staging_ready) {
        p->staging = p->shape;
        for (int l = 0; l < p->shape.n_levels; ++l) {
            const size_t hw = static_cast<size_t>(p->shape.lv[l].H) * p->shape.lv[l].W;
            float *a = nullptr, *b = nullptr;
            DAS_TRY(dev_alloc(&a, B * hw));
            DAS_TRY(dev_alloc(&b, B * hw));
            p->staging.lv[l].cls = a; p->staging.lv[l].ctr = b; p->staging.lv[l].pose = nullptr;
            for (int k = 0; k < DAS_MAX_LAYERS; ++k) p->staging.lv[l].feats[k] = nullptr;
        }
        p->staging_ready = true;
    }
    DAS_TRY(set_row_cache(p, zero_copy && p->host_mode == 2));
    das_levels run = p->staging;
    run.in_dtype = levels->in_dtype;
    const size_t es = static_cast<size_t>(dtype_bytes(levels->in_dtype));     // element size of cls / ctr / pose
    int64_t copied = 0;
    for (int l = 0; l < p->shape.n_levels; ++l) {
        const das_level_desc& h = levels->lv[l];
        das_level_desc& d = run.lv[l];
        DAS_REQUIRE(h.H == d.H && h.W == d.W && h.stride == d.stride, DAS_ERR_ARG, "run_host: level %d shape differs", l);
        DAS_REQUIRE(h.cls && h.ctr && h.pose, DAS_ERR_ARG, "run_host: level %d has a null map", l);
        const size_t hw = static_cast<size_t>(h.H) * h.W;
        d.scale_offset = h.scale_offset; d.scale_depth = h.scale_depth; d.scale_uv = h.scale_uv; d.scale_d = h.scale_d;
        DAS_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(d.cls), h.cls, B * hw * es, cudaMemcpyHostToDevice, st));
        copied += static_cast<int64_t>(B * hw * es);
        // The bounded fast path of score_topk (no peak mask, nms_pre <= 128) sweeps the cls plane only and looks the
        // centerness up at a few dozen cells per image: in the zero-copy modes that plane stays in pinned host memory too
        // (6.8 MB less over PCIe per 64-image batch); otherwise every centerness value is needed and the plane is copied.
        const float* ctr_alias = nullptr;
        if (zero_copy && p->cfg.peak_kernel != 3 && level_slots(static_cast<int>(hw), p->cfg.nms_pre) <= 128 &&
            level_slots(static_cast<int>(hw), p->cfg.nms_pre) < static_cast<int>(hw)) {
            cudaPointerAttributes at{};
            if (cudaPointerGetAttributes(&at, h.ctr) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
                ctr_alias = static_cast<const float*>(at.devicePointer);
            else cudaGetLastError();
        }
        if (ctr_alias) {
            d.ctr = ctr_alias;
        } else {
            DAS_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(d.ctr), h.ctr, B * hw * es, cudaMemcpyHostToDevice, st));
            copied += static_cast<int64_t>(B * hw * es);
        }
        das_level_desc& sd = p->staging.lv[l];      // bulk staging, allocated lazily
        if (zero_copy && pose_sparse) {
            d.pose = dev_alias[l][0];
        } else {
            if (!sd.pose) { float* c = nullptr; DAS_TRY(dev_alloc(&c, B * hw * (3 + 6 * J))); sd.pose = c; }
            DAS_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(sd.pose), h.pose, B * hw * (3 + 6 * J) * es, cudaMemcpyHostToDevice, st));
            copied += static_cast<int64_t>(B * hw * (3 + 6 * J) * es);
            d.pose = sd.pose;
        }
        for (int k = 0; k < L; ++k) {
            DAS_REQUIRE(h.feats[k], DAS_ERR_ARG, "run_host: level %d layer %d feature map is null", l, k);
            if (zero_copy && k == L - 1) {          // the last layer is evaluated sparsely at the selected centres
                d.feats[k] = dev_alias[l][1 + k];
                continue;
            }
            if (!sd.feats[k]) { float* f = nullptr; DAS_TRY(dev_alloc(&f, B * hw * C)); sd.feats[k] = f; }
            DAS_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(sd.feats[k]), h.feats[k], B * hw * C * 4, cudaMemcpyHostToDevice, st));
            copied += static_cast<int64_t>(B * hw * C * 4);
            d.feats[k] = sd.feats[k];
        }
    }
    p->h2d_explicit = copied + static_cast<int64_t>(B) * (2 * 4 + DAS_CAM_DOUBLES * 8);
    p->in_host_call = true;
    const int bound = das_plan_bind(p, &run, st);
    p->in_host_call = false;
    DAS_TRY(bound);
    DAS_TRY(das_plan_set_metas(p, scale_xy, cam, st));
    DAS_TRY(das_plan_run(p, st, 1));
    const size_t P = p->P;
    auto D2H = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess;
    };
    // host_out laid out like the device block (das_plan_output_block; DecodePlan.alloc_host_out does that): one copy
    // instead of seven (each small copy costs the stream ~5-8 us)
    {
        unsigned char* hb = reinterpret_cast<unsigned char*>(host_out.out_count);
        const void* hp[7] = {host_out.out_count, host_out.out_score, host_out.out_slot, host_out.out_pose, host_out.out_center,
                             host_out.out_cam, host_out.out_world};
        bool block_layout = hb != nullptr && p->out_off[0] == 0;
        for (int i = 0; i < 7 && block_layout; ++i) block_layout = hp[i] == hb + p->out_off[i];
        if (block_layout) {
            DAS_CUDA_CHECK(cudaMemcpyAsync(hb, p->out_block, p->out_block_bytes, cudaMemcpyDeviceToHost, st));
        } else {
            DAS_CUDA_CHECK(D2H(host_out.out_count, p->buf.out_count, B * 4));
            DAS_CUDA_CHECK(D2H(host_out.out_score, p->buf.out_score, B * P * 4));
            DAS_CUDA_CHECK(D2H(host_out.out_slot, p->buf.out_slot, B * P * 4));
            DAS_CUDA_CHECK(D2H(host_out.out_pose, p->buf.out_pose, B * P * J * 3 * 4));
            DAS_CUDA_CHECK(D2H(host_out.out_center, p->buf.out_center, B * P * 3 * 4));
            DAS_CUDA_CHECK(D2H(host_out.out_cam, p->buf.out_cam, B * P * J * 3 * 8));
            DAS_CUDA_CHECK(D2H(host_out.out_world, p->buf.out_world, B * P * J * 3 * 8));
        }
    }
    DAS_CUDA_CHECK(D2H(host_out.cand_score, p->buf.cand_score, B * p->CT * 4));
    DAS_CUDA_CHECK(D2H(host_out.cand_index, p->buf.cand_index, B * p->CT * 4));
    return DAS_OK;
}
