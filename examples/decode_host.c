/* A C99 host binding the decode through include/das_decode.h alone -- what a non-Python caller of DASHead.get_poses
 * (reference: mmdet3d/models/pose_heads/das_head.py:653-688) links against.  It builds a plan for BASELINE config #1
 * (one Panoptic-shaped image: J = 15, 128x208 map, stride 8, K = 10) and reports the plan's static shapes.  The inputs of a
 * real decode are DEVICE pointers produced by the network in front of the path (das_plan_bind), or pinned HOST pointers through
 * das_plan_run_host; this example owns neither, so it stops after the plan.  There is no CPU path: without a CUDA device
 * das_plan_create fails and the program prints the library's message and exits with status 2.
 *
 *   gcc -std=c99 -Wall -Iinclude examples/decode_host.c -Ldas_b200/lib -ldas_decode -Wl,-rpath,$PWD/das_b200/lib -o decode_host
 */
#include <stdio.h>
#include <string.h>

#include "das_decode.h"

int main(void) {
    das_decode_cfg cfg;
    das_levels shape;
    das_plan* plan = NULL;
    das_buffers out;
    int32_t cand_slots = 0, out_slots = 0;
    void* block = NULL;
    int64_t block_bytes = 0;
    int st;

    memset(&cfg, 0, sizeof cfg);
    memset(&shape, 0, sizeof shape);
    cfg.num_joints = 15;           /* configs/das/exp_panoptic.py:7,38-40 */
    cfg.root_idx = 2;
    cfg.num_heads = 4;             /* configs/_base_/models/das.py:43-50 */
    cfg.feat_channels = 256;
    cfg.num_layers = 1;
    cfg.depth_factor = 20.f;
    cfg.z_norm = 50.f;
    cfg.nms_pre = 10;              /* test_cfg: K = nms_pre = nms_post */
    cfg.nms_post = 10;
    cfg.nms_thr = 0.9f;
    cfg.score_thr = 0.f;
    cfg.refine = 1;                /* pose maps are the raw predictor output; refinement + eval tail run in the library */
    cfg.dataset_depth_factor = 1.0;
    shape.n_levels = 1;
    shape.batch = 1;
    shape.in_dtype = DAS_DTYPE_F32;
    shape.lv[0].H = 128;
    shape.lv[0].W = 208;
    shape.lv[0].stride = 8;

    printf("%s\n", das_version());
    printf("candidate slots per image: %d, output slots: %d\n", (int)das_candidate_slots(&shape, cfg.nms_pre),
           (int)das_output_slots(das_candidate_slots(&shape, cfg.nms_pre), cfg.nms_post));
    st = das_plan_create(&cfg, &shape, &plan);
    if (st != DAS_OK) {
        fprintf(stderr, "das_plan_create failed (%d): %s\n", st, das_last_error());
        return 2;
    }
    st = das_plan_buffers(plan, &out, &cand_slots, &out_slots);
    if (st == DAS_OK) st = das_plan_output_block(plan, &block, &block_bytes);
    if (st != DAS_OK) {
        fprintf(stderr, "plan query failed (%d): %s\n", st, das_last_error());
        das_plan_destroy(plan);
        return 2;
    }
    printf("plan ready: %d candidate slots, %d output slots, packed result block of %lld bytes at %p\n", (int)cand_slots,
           (int)out_slots, (long long)block_bytes, block);
    das_plan_destroy(plan);
    return 0;
}
