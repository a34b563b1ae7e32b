"""GPU: the CUDA path, called through the C-ABI plan, against the CPU oracle on the same seeded inputs,
against the golden vectors of the reference's own code, and through size-independent properties at
the full BASELINE size.

Bars: person order / selected cells BIT-EXACT; scores within 8 ulp; 3D joint coordinates
within 1e-4 relative to their magnitude (floor 1 unit).  Every case first ASSERTS that it is decidable -- rank
margins >= 16 ulp over the candidates that survive score_thr (the CPU reference's own fp32 sigmoid is only accurate
to ~2 ulp) and every OKS decision >= 1e-6 from nms_thr; the generator reject-samples to
guarantee it (synth.enforce_rank_margins) -- so no comparison is ever skipped.
"""
import dataclasses
import os

import numpy as np
import pytest
import torch

import util
from das_b200 import synth
from das_b200.head import DASHeadB200
from oracle import das_oracle as O
from oracle import make_golden as G

pytestmark = pytest.mark.gpu

P = synth.PANOPTIC
TOL = 1e-4
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def close(g, o, e=None):
    """|g - o| <= TOL relative to magnitude (floor 1 unit).  With an fp64 arbiter result `e` (multi-layer refinement at
    full map size, where three chained layers amplify fp32 round-off) the value also passes when the GPU is as close to
    the exact answer as the fp32 reference itself is: err(g, e) <= 3 * err(o, e) + 1e-5 -- two fp32 implementations cannot
    agree better than either agrees with the truth."""
    g, o = np.asarray(g, dtype=np.float64), np.asarray(o, dtype=np.float64)
    if util.rel_err(g, o) < TOL:
        return True
    if e is None:
        return False
    e = np.asarray(e, dtype=np.float64)
    return util.rel_err(g, o) < 3 * TOL and util.rel_err(g, e) <= 3 * util.rel_err(o, e) + 1e-5


def compare(plan, got, ref, ref64=None):
    """Person order and source cells bit-exact, values within the bars.  Callers assert the case's margins first
    (util.assert_margins), so there is no tolerance on the order and nothing is skipped."""
    ci = plan.t["cand_index"].cpu()
    assert len(got) == len(ref)
    for b, (g, o) in enumerate(zip(got, ref)):
        e = ref64[b] if ref64 is not None else None
        if e is not None:
            assert e["index"].tolist() == o["index"].tolist()
        lv, idx = util.slot_to_level_index(plan, g["slots"].cpu(), ci[b])
        ref_pairs = list(zip(o["level"].tolist(), o["index"].tolist()))
        assert list(zip(lv, idx)) == ref_pairs, f"image {b}: person order / source cells differ"
        if not idx:
            assert g["poses"].shape[0] == 0 and g["scores"] == []
            continue
        assert int(util.ulp_gap(torch.tensor(g["scores"]), o["scores"]).max()) <= 8
        assert close(g["poses"].cpu().numpy(), o["poses"].numpy(), e and e["poses"].numpy())
        assert close(g["centers"].cpu().numpy(), o["centers"].numpy(), e and e["centers"].numpy())
        assert close(g["poses_cam"].cpu().numpy(), o["poses_cam"], e and e["poses_cam"])
        assert close(g["poses_world"].cpu().numpy(), o["poses_world"], e and e["poses_world"])
        assert torch.all(g["vis"] == 1)


CASES = [
    # id, cfg, B, H, W, test_cfg, kwargs
    ("small", P, 2, 24, 40, dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0), {}),
    ("cfg1_128x208", P, 1, 128, 208, dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0), {}),
    ("scales_thr", P, 2, 64, 96, dict(nms_pre=50, nms_post=20, nms_thr=0.9, score_thr=0.07),
     dict(scales=(1.1, 0.9, 1.05, 0.95), peaks=24)),
    ("pyramid4", dataclasses.replace(P, strides=(8, 16, 32, 64)), 2, 64, 96,
     dict(nms_pre=100, nms_post=30, nms_thr=0.9, score_thr=0.05), dict(peaks=24)),
    ("mupots17", dataclasses.replace(synth.MUPOTS17, num_layers=1), 2, 48, 64,
     dict(nms_pre=20, nms_post=20, nms_thr=0.9, score_thr=0.0), {}),
    ("coherent_nms", P, 2, 48, 64, dict(nms_pre=30, nms_post=30, nms_thr=0.9, score_thr=0.0), dict(coherent=8)),
    ("reference_cfg", P, 1, 80, 144, dict(nms_pre=1000, nms_post=100, nms_thr=0.9, score_thr=0.07), dict(peaks=30)),
    ("crowded_cfg4", P, 1, 256, 416, dict(nms_pre=64, nms_post=64, nms_thr=0.9, score_thr=0.0), dict(peaks=80)),
    ("no_nms", P, 2, 24, 40, dict(nms_pre=12, nms_thr=0.9, score_thr=0.02), {}),
    ("odd_size_unaligned", P, 3, 23, 37, dict(nms_pre=9, nms_post=9, nms_thr=0.9, score_thr=0.0), {}),
]


@pytest.mark.parametrize("case_id,cfg,B,H,W,tc,kw", CASES, ids=[c[0] for c in CASES])
def test_full_path_matches_oracle(case_id, cfg, B, H, W, tc, kw):
    case = util.make_case(cfg, B, H, W, seed=1234, tc=tc, **kw)
    ref, _ = util.run_oracle(case, tc)
    util.assert_margins(case, tc, ref)
    plan, got = util.run_gpu(case, tc, refine=True)
    compare(plan, got, ref)


@pytest.mark.parametrize("case_id,cfg,B,H,W,tc,kw", CASES[:5], ids=[c[0] for c in CASES[:5]])
def test_reference_contract_decode_only(case_id, cfg, B, H, W, tc, kw):
    """get_poses with already-refined pose maps (the reference signature): the decode arithmetic is the same
    fp32 op sequence, so poses must be BIT-equal to the oracle's."""
    case = util.make_case(cfg, B, H, W, seed=99, tc=tc, **kw)
    ref, pose_preds = util.run_oracle(case, tc)
    util.assert_margins(case, tc, ref)
    plan, got = util.run_gpu(case, tc, refine=False, pose_override=pose_preds)
    compare(plan, got, ref)
    for g, o in zip(got, ref):
        assert torch.equal(g["poses"].cpu(), o["poses"]) and torch.equal(g["centers"].cpu(), o["centers"])


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_golden_vectors_of_the_reference(name):
    if G.CASES[name][0].num_layers > 1:
        pytest.skip("num_layers > 1 is covered by test_gpu_dense_layers.py")
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg, levels, layers, metas, tc = G.build_case(name)
    np.testing.assert_allclose(G.checksum(levels), gold["checksum"], rtol=1e-9, atol=1e-6)
    case = dict(cfg=cfg, levels=levels, layers=layers, metas=metas, batch=levels[0]["cls"].shape[0])
    # the golden case is decidable (margins were asserted and stored when it was generated; re-checked here)
    assert util.assert_margins(levels, tc) == int(gold["rank_margin_ulps"]) and float(gold["oks_margin"]) >= util.MIN_OKS_MARGIN
    plan, got = util.run_gpu(case, tc, refine=True)
    ci = plan.t["cand_index"].cpu()
    assert len(got) == int(gold["n_images"])
    for i, g in enumerate(got):
        lv, idx = util.slot_to_level_index(plan, g["slots"].cpu(), ci[i])
        assert idx == gold[f"index_{i}"].tolist() and lv == gold[f"level_{i}"].tolist()
        if idx:
            assert int(util.ulp_gap(torch.tensor(g["scores"], dtype=torch.float32), torch.tensor(gold[f"scores_{i}"])).max()) <= 8
        assert util.rel_err(g["poses"].cpu().numpy(), gold[f"poses_{i}"]) < TOL
        assert util.rel_err(g["poses_cam"].cpu().numpy(), gold[f"cam_{i}"]) < TOL
        assert util.rel_err(g["poses_world"].cpu().numpy(), gold[f"world_{i}"]) < TOL


def test_peak_mask_mode_matches_oracle_variant():
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    for (h, w) in ((48, 64), (130, 210)):
        case = util.make_case(P, 2, h, w, seed=5, tc=tc, peak_kernel=3)
        ref, _ = util.run_oracle(case, tc, peak_kernel=3)
        util.assert_margins(case, tc, ref, peak_kernel=3)
        plan, got = util.run_gpu(case, tc, refine=True, peak_kernel=3)
        compare(plan, got, ref)
        # a 3x3 bump contributes exactly one candidate under the peak mask
        for g in got:
            assert len(set(g["slots"].tolist())) == len(g["scores"])


def test_ties_break_towards_lower_index():
    """Constant logits: every score ties. Rule: lower cell index first (north_star), pinned by the stable oracle."""
    tc = dict(nms_pre=7, nms_thr=0.9, score_thr=0.0)         # no nms_post: order = top-k order
    case = util.make_case(P, 2, 16, 20, seed=3)
    for lv in case["levels"]:
        lv["cls"].fill_(0.25)
        lv["ctr"].fill_(-0.5)
        lv["cls"][0, 0, 5, 7] = 3.0                           # one clear winner, then ties
    ref, _ = util.run_oracle(case, tc, stable=True)
    plan, got = util.run_gpu(case, tc, refine=True)
    ci = plan.t["cand_index"].cpu()
    assert ci[0, :7].tolist() == [5 * 20 + 7, 0, 1, 2, 3, 4, 5]
    assert ci[1, :7].tolist() == [0, 1, 2, 3, 4, 5, 6]
    compare(plan, got, ref)          # exact ties by construction: the order is the tie rule's, nothing to excuse


def test_large_k_exact_select_with_many_ties():
    """K > 128 takes the exact radix path; quantised logits create heavy ties at the K-th score."""
    tc = dict(nms_pre=300, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 2, 40, 50, seed=8)
    for lv in case["levels"]:
        lv["cls"].copy_(torch.round(lv["cls"] * 2) / 2)
        lv["ctr"].fill_(0.0)
    ref, _ = util.run_oracle(case, tc, stable=True)
    plan, got = util.run_gpu(case, tc, refine=True)
    ci = plan.t["cand_index"].cpu()
    for b, o in enumerate(ref):
        assert ci[b].tolist() == o["cand_index"].tolist()


def test_nothing_survives_score_thr():
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.9999)
    case = util.make_case(P, 2, 24, 40, seed=11)
    plan, got = util.run_gpu(case, tc, refine=True)
    for g in got:
        assert g["poses"].shape == (0, 15, 3) and g["scores"] == [] and g["centers"].shape == (0, 3)


def test_pass_through_level_keeps_raster_order():
    """HW <= nms_pre: no top-k, every cell in raster order (das_head.py:717)."""
    tc = dict(nms_pre=100, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 1, 6, 9, seed=12)
    ref, _ = util.run_oracle(case, tc)
    plan, got = util.run_gpu(case, tc, refine=True)
    assert plan.t["cand_index"].cpu()[0].tolist() == list(range(54))
    compare(plan, got, ref)          # no nms_post: output order = raster order, independent of the scores


def test_targets_outside_the_map_sample_zero():
    """Huge offsets push every sampling location out of the map: grid_sample's zero padding."""
    tc = dict(nms_pre=8, nms_post=8, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 2, 20, 28, seed=13, tc=tc)
    J = P.num_joints
    for lv in case["levels"]:
        lv["pose_raw"][:, 3:3 + 3 * J:3] += 500.0
        lv["pose_raw"][1, 4:3 + 3 * J:3] -= 37.25      # and some joints only partially outside
    ref, _ = util.run_oracle(case, tc)
    util.assert_margins(case, tc, ref)
    plan, got = util.run_gpu(case, tc, refine=True)
    compare(plan, got, ref)


def test_white_noise_fields_within_reference_noise_floor():
    """Unsmoothed pose fields amplify fp32 coordinate round-off; the fp64 run of the same algorithm arbitrates:
    the GPU must be as close to it as the fp32 reference is, within 3x + 1e-5."""
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 2, 32, 48, seed=21, smooth=1, tc=tc)
    ref32, _ = util.run_oracle(case, tc)
    util.assert_margins(case, tc, ref32)
    ref64 = util.run_oracle64(case["levels"], case["layers"], case["metas"], P, tc)
    plan, got = util.run_gpu(case, tc, refine=True)
    ci = plan.t["cand_index"].cpu()
    for b, (g, a, e) in enumerate(zip(got, ref32, ref64)):
        # scores (hence the candidates) do not depend on the pose arithmetic, and random poses never overlap in OKS
        assert a["index"].tolist() == e["index"].tolist()
        assert util.slot_to_level_index(plan, g["slots"].cpu(), ci[b])[1] == a["index"].tolist()
        err_ref = util.rel_err(a["poses"].numpy(), e["poses"].numpy())
        err_gpu = util.rel_err(g["poses"].cpu().numpy(), e["poses"].numpy())
        assert err_gpu <= 3 * err_ref + 1e-5, (err_gpu, err_ref)


def test_graph_replay_eager_and_repeat_are_identical():
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 4, 32, 48, seed=31)
    plan, got = util.run_gpu(case, tc, refine=True, use_graph=False)
    a = plan.output_block().clone()
    for _ in range(3):
        plan.run(use_graph=True)
    torch.cuda.synchronize()
    b = plan.output_block().clone()
    plan.run(stage_events=True)
    torch.cuda.synchronize()
    c = plan.output_block().clone()
    assert torch.equal(a, b) and torch.equal(a, c)
    assert all(t >= 0 for t in plan.stage_ms())
    assert plan.kernel_launches >= 3 * 5


@pytest.mark.parametrize("shape", [(3, 32, 48, 1), (24, 64, 96, 1), (2, 24, 40, 3)], ids=["small", "large", "three_layers"])
def test_programmatic_dependent_launch_chain_is_bit_identical(shape):
    """das_plan_set_pdl: with programmatic dependent launch every kernel of the chain is scheduled while its predecessor
    still runs and waits for it on the device (griddepcontrol.wait); the counters the refinement uses are cleared by the
    top-k kernel instead of memset nodes.  Eager, graph and repeated replays must equal the plain stream-ordered chain bit
    for bit -- also for the batched phase 1-2 kernel (large) and with dense layers in between (three_layers)."""
    B, h, w, L = shape
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    cfg = dataclasses.replace(P, num_layers=L)
    case = util.make_case(cfg, B, h, w, seed=33)
    blocks = {}
    for mode in (0, 1):
        plan = util.make_plan(case, tc)
        plan.set_pdl(mode)
        dl = synth.levels_to(case["levels"], "cuda")
        plan.bind([dict(cls=lv["cls"], ctr=lv["ctr"], pose=lv["pose_raw"], feats=lv["feats"], scales=lv["scales"]) for lv in dl])
        plan.set_metas(case["metas"])
        plan.run(use_graph=False)
        torch.cuda.synchronize()
        eager = plan.output_block().clone()
        for _ in range(4):
            plan.run(use_graph=True)
        torch.cuda.synchronize()
        assert torch.equal(eager, plan.output_block()), f"pdl={mode}: graph replay differs from the eager run"
        blocks[mode] = eager
    assert torch.equal(blocks[0], blocks[1])
    assert int(plan.views_of_block(blocks[1])["out_count"].sum()) > 0


def test_host_entry_equals_device_entry():
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 3, 24, 40, seed=41, scales=(1.1, 0.9, 1.05, 0.95))
    plan, got = util.run_gpu(case, tc, refine=True)
    host_levels = [dict(cls=lv["cls"].pin_memory(), ctr=lv["ctr"].pin_memory(), pose=lv["pose_raw"].pin_memory(),
                        feats=[f.permute(0, 2, 3, 1).contiguous().pin_memory().permute(0, 3, 1, 2) for f in lv["feats"]],
                        scales=lv["scales"]) for lv in case["levels"]]
    out = plan.alloc_host_out(contiguous=False)          # one host array per field: seven D2H copies
    plan.run_host(host_levels, case["metas"], out)
    res = plan.results(case["metas"], src=out)
    for g, h in zip(got, res):
        assert g["scores"] == h["scores"] and torch.equal(g["poses"].cpu(), h["poses"]) and torch.equal(g["poses_cam"].cpu(), h["poses_cam"])
    assert plan.h2d_bytes > 3 * 24 * 40 * 256 * 4 and plan.d2h_bytes > 0
    assert plan.h2d_explicit_bytes == plan.h2d_bytes
    # zero-copy policy: only the logit planes are copied, pose / feature maps are read in place from pinned memory
    plan.set_host_mode(True)
    out2 = plan.alloc_host_out()
    plan.run_host(host_levels, case["metas"], out2)
    assert plan.h2d_explicit_bytes < plan.h2d_bytes // 20
    for k in out:
        assert torch.equal(out[k], out2[k]), k
    # ... without the row cache, with a row cache too small for the distinct rows, and back to the cached default:
    # the same bits every time (a full cache leaves the remaining records pointing at the host rows)
    for row_cache, cap in ((False, None), (True, "7"), (True, None)):
        if cap is not None:
            os.environ["DAS_ROW_CACHE_ROWS"] = cap
        p2 = util.make_plan(case, tc, refine=True)
        os.environ.pop("DAS_ROW_CACHE_ROWS", None)
        p2.set_host_mode(True, row_cache=row_cache)
        o = p2.alloc_host_out()
        for _ in range(3):                     # eager first run, then graph capture + replay
            for t in o.values():
                t.zero_()
            p2.run_host(host_levels, case["metas"], o)
            for k in out:
                assert torch.equal(out[k], o[k]), (row_cache, cap, k)
    # pageable inputs silently fall back to staging copies
    pageable = [dict(cls=lv["cls"].clone(), ctr=lv["ctr"].clone(), pose=lv["pose_raw"].clone(),
                     feats=[f.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2) for f in lv["feats"]], scales=lv["scales"])
                for lv in case["levels"]]
    out3 = plan.alloc_host_out()
    plan.run_host(pageable, case["metas"], out3)
    assert plan.h2d_explicit_bytes == plan.h2d_bytes
    for k in out:
        assert torch.equal(out[k], out3[k]), k


def test_async_host_entry_two_plans_in_flight():
    """das_plan_run_host_async: two plans on two streams driven in turn (the bench's e2e loop) deliver the bits of the
    synchronous call, in every transfer policy."""
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    cases = [util.make_case(P, 4, 48, 64, seed=43 + i, scales=(1.1, 0.9, 1.05, 0.95)) for i in range(2)]
    host = [[dict(cls=lv["cls"].pin_memory(), ctr=lv["ctr"].pin_memory(), pose=lv["pose_raw"].pin_memory(),
                  feats=[f.permute(0, 2, 3, 1).contiguous().pin_memory().permute(0, 3, 1, 2) for f in lv["feats"]],
                  scales=lv["scales"]) for lv in c["levels"]] for c in cases]
    want = []
    for c, h in zip(cases, host):
        plan = util.make_plan(c, tc, refine=True)
        o = plan.alloc_host_out()
        plan.run_host(h, c["metas"], o)
        assert int(o["out_count"].sum()) > 0
        want.append({k: v.clone() for k, v in o.items()})
    for zero_copy, row_cache in ((False, False), (True, False), (True, True)):
        plans = [util.make_plan(c, tc, refine=True) for c in cases]
        outs = [p.alloc_host_out() for p in plans]
        streams = [torch.cuda.Stream() for _ in plans]
        for p in plans:
            p.set_host_mode(zero_copy, row_cache=row_cache)
        for i in range(8):
            k = i % 2
            streams[k].synchronize()
            if i >= 2:
                for key, v in want[k].items():
                    assert torch.equal(outs[k][key], v), (zero_copy, row_cache, i, key)
                for t in outs[k].values():
                    t.zero_()
            with torch.cuda.stream(streams[k]):
                plans[k].run_host(host[k], cases[k]["metas"], outs[k], sync=False)
        torch.cuda.synchronize()
        for k in range(2):
            for key, v in want[k].items():
                assert torch.equal(outs[k][key], v), (zero_copy, row_cache, "last", key)


@pytest.mark.parametrize("cfg,tc,kw", [
    (P, dict(nms_pre=30, nms_post=30, nms_thr=0.9, score_thr=0.0), dict(coherent=8, peaks=6)),
    (dataclasses.replace(P, strides=(8, 16, 32, 64)), dict(nms_pre=40, nms_post=30, nms_thr=0.9, score_thr=0.05), dict(peaks=24)),
], ids=["coherent", "pyramid4_score_thr"])
def test_host_row_cache_with_heavy_row_sharing(cfg, tc, kw):
    """Coherent fields: neighbouring candidates point at the same joints, so many warps ask for the same feature rows at
    the same time -- the first fetches, the others wait for its copy.  Must stay bit-equal to the device entry
    (also on a 4-level pyramid with candidates dropped by score_thr)."""
    case = util.make_case(cfg, 8, 64, 96, seed=77, **kw)
    plan, _ = util.run_gpu(case, tc, refine=True)
    want = plan.output_block().clone()
    host_levels = [dict(cls=lv["cls"].pin_memory(), ctr=lv["ctr"].pin_memory(), pose=lv["pose_raw"].pin_memory(),
                        feats=[f.permute(0, 2, 3, 1).contiguous().pin_memory().permute(0, 3, 1, 2) for f in lv["feats"]],
                        scales=lv["scales"]) for lv in case["levels"]]
    plan.set_host_mode(True)
    out = plan.alloc_host_out()
    ref = plan.views_of_block(want.cpu())
    for _ in range(6):
        plan.run_host(host_levels, case["metas"], out)
        for k, v in ref.items():
            assert torch.equal(out[k], v), k


def test_drop_in_head_api():
    """DASHeadB200.get_poses with the reference signature and return structure (das_head.py:653-688)."""
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 2, 24, 40, seed=51)
    ref, pose_preds = util.run_oracle(case, tc)
    head = DASHeadB200(num_classes=1, in_channels=256, num_joints=15, strides=[8], depth_factor=20, z_norm=50, root_idx=2,
                       recursive_update=dict(num_heads=4, feat_channels=256, num_layers=1, dim=3, num_joints=15), test_cfg=tc,
                       loss_cls=dict(type="FocalLoss"), train_cfg=None)       # unrelated keys are accepted
    dl = synth.levels_to(case["levels"], "cuda")
    # (a) reference call: refined maps
    res = head.get_poses([lv["cls"] for lv in dl], [p.cuda() for p in pose_preds], [lv["ctr"] for lv in dl], case["metas"],
                         rescale=True)
    assert isinstance(res, list) and len(res) == 2
    for r, o, m in zip(res, ref, case["metas"]):
        assert set(r) >= {"poses", "vis", "centers", "image_paths", "scores", "poses_cam"}
        assert isinstance(r["scores"], list) and r["image_paths"] == [m["filename"]]
        assert r["poses"].is_cuda and tuple(r["poses"].shape) == tuple(o["poses"].shape) and r["vis"].shape == r["poses"].shape[:2]
        assert util.rel_err(r["poses"].cpu().numpy(), o["poses"].numpy()) < TOL
    # (b) extended call: raw maps + refinement features
    head.load_refine_weights(synth.layers_to(case["layers"], "cuda"))
    res2 = head.get_poses([lv["cls"] for lv in dl], [lv["pose_raw"] for lv in dl], [lv["ctr"] for lv in dl],
                          [lv["feats"] for lv in dl], case["metas"])
    for r, o in zip(res2, ref):
        assert util.rel_err(r["poses_cam"].cpu().numpy(), o["poses_cam"]) < TOL
    # (c) the reference's assert on mismatched level lists
    with pytest.raises(AssertionError):
        head.get_poses([dl[0]["cls"]], [], [dl[0]["ctr"]], case["metas"])


def test_full_size_config2_all_images_and_properties():
    """BASELINE config #2 at full size (B=64, J=15, 128x208, K=10, L=1), generated on the device: EVERY image against the
    CPU oracle (run on host copies in chunks), plus the size-independent properties the domain offers."""
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    B, H, W = 64, 128, 208
    case, plan, got, ref = util.run_full_size(P, B, H, W, tc, seed=1234, peaks=16, chunk=8)
    util.assert_margins(case, tc, ref)
    compare(plan, got, ref)
    t = {k: v.clone() for k, v in plan.t.items()}
    lv = case["levels"][0]
    # selected cells are exactly torch.topk's on the same device scores, in the same order
    score = (lv["cls"].sigmoid() * lv["ctr"].sigmoid()).flatten(1)
    top_v, top_i = score.topk(10, dim=1)
    assert torch.equal(t["cand_index"].long(), top_i)
    assert torch.all(t["cand_score"][:, :-1] >= t["cand_score"][:, 1:])
    assert int(util.ulp_gap(t["cand_score"].cpu(), top_v.cpu()).max()) <= 8
    # counts, ordering and finiteness of the outputs
    cnt = t["out_count"]
    assert int(cnt.min()) >= 1 and int(cnt.max()) <= 10
    for b in range(B):
        n = int(cnt[b])
        s = t["out_score"][b, :n]
        assert torch.all(s[:-1] >= s[1:])
        assert torch.isfinite(t["out_pose"][b, :n]).all() and torch.isfinite(t["out_cam"][b, :n]).all()
        assert torch.all(t["out_score"][b, n:] == 0)
    # batch independence: image 5 decoded alone gives the same bits (images are independent, das_head.py:666)
    one = dict(cfg=P, levels=[dict(lv, cls=lv["cls"][5:6], ctr=lv["ctr"][5:6], pose_raw=lv["pose_raw"][5:6],
                                   feats=[f[5:6] for f in lv["feats"]])], layers=case["layers"], metas=case["metas"][5:6], batch=1)
    p1, _ = util.run_gpu(one, tc, refine=True)
    assert torch.equal(p1.t["out_pose"][0], t["out_pose"][5]) and torch.equal(p1.t["out_cam"][0], t["out_cam"][5])


def test_full_size_config3_mupots_three_layers():
    """BASELINE config #3 at its real map size: J=17 (COCO sigma table in OKS), L=3 (two dense layers + the sparse one),
    K=20, 128x208; a 4-image batch, every image against the oracle."""
    tc = dict(nms_pre=20, nms_post=20, nms_thr=0.9, score_thr=0.0)
    case, plan, got, ref, ref64 = util.run_full_size(synth.MUPOTS17, 4, 128, 208, tc, seed=1239, peaks=30, chunk=2, arbiter=True)
    util.assert_margins(case, tc, ref)
    compare(plan, got, ref, ref64)          # three chained layers at full size: judged against the fp64 arbiter (see close())


def test_full_size_config4_crowded():
    """BASELINE config #4: 256x416 map, K=64 people per image; a 3-image batch, every image against the oracle."""
    tc = dict(nms_pre=64, nms_post=64, nms_thr=0.9, score_thr=0.0)
    case, plan, got, ref = util.run_full_size(P, 3, 256, 416, tc, seed=1240, peaks=96, chunk=1)
    util.assert_margins(case, tc, ref)
    compare(plan, got, ref)
    assert all(len(g["scores"]) == 64 for g in got)


@pytest.mark.parametrize("mode,tol", [(1, TOL), (2, 5e-2)])
def test_tensor_core_refinement_modes(mode, tol):
    """tcgen05 path (das_refine_heads + das_refine_tc): 3xTF32 meets the fp32 bar, single-pass TF32 is looser."""
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    for (B, H, W, seed, kw) in ((3, 40, 56, 71, {}), (2, 64, 96, 72, dict(scales=(1.1, 0.9, 1.05, 0.95)))):
        case = util.make_case(P, B, H, W, seed=seed, tc=tc, **kw)
        ref, _ = util.run_oracle(case, tc)
        util.assert_margins(case, tc, ref)
        plan, got = util.run_gpu(case, tc, refine=True, refine_mode=mode)
        plan0, got0 = util.run_gpu(case, tc, refine=True, refine_mode=0)
        for g, g0, o in zip(got, got0, ref):
            assert g["scores"] == g0["scores"] and g["slots"].tolist() == g0["slots"].tolist()
            assert util.rel_err(g["poses"].cpu().numpy(), o["poses"].numpy()) < tol
            assert util.rel_err(g["poses_cam"].cpu().numpy(), o["poses_cam"]) < tol


@pytest.fixture
def heads_kernel(request):
    """Forces the phase 1-2 kernel das_refine_heads launches (1 = warp per item, 2 = 4 candidates per warp); small test
    decodes would otherwise only ever see the warp-per-item kernel."""
    from das_b200 import _lib
    lib = _lib.load()
    assert lib.das_debug_force_heads_kernel(request.param) == 0
    yield request.param
    lib.das_debug_force_heads_kernel(0)


@pytest.mark.parametrize("heads_kernel", [1, 2], indirect=True, ids=["per_item", "batched"])
def test_tensor_core_refinement_with_threshold_and_pyramid(heads_kernel):
    cfg = dataclasses.replace(P, strides=(8, 16, 32))
    tc = dict(nms_pre=60, nms_post=30, nms_thr=0.9, score_thr=0.05)
    case = util.make_case(cfg, 2, 48, 64, seed=73, peaks=20, tc=tc)
    ref, _ = util.run_oracle(case, tc)
    util.assert_margins(case, tc, ref)
    plan, got = util.run_gpu(case, tc, refine=True, refine_mode=1)
    compare(plan, got, ref)


@pytest.mark.parametrize("heads_kernel", [1, 2], indirect=True, ids=["per_item", "batched"])
@pytest.mark.parametrize("name", ["odd_sizes_3_images", "border_targets", "three_layers_j17"])
def test_both_phase12_kernels_match_the_oracle(heads_kernel, name):
    """The batched phase 1-2 kernel (4 candidates per warp, a CTA per joint, weights in shared memory, one row-list
    reservation per task) against the oracle on the shapes that stress its bookkeeping: a candidate count that is not a
    multiple of 4 on an odd-sized map, sampling targets outside the map, and dense layers in front (previous offsets from
    the joint-major maps, J = 17)."""
    if name == "odd_sizes_3_images":
        cfg, B, h, w, tc, kw = P, 3, 23, 37, dict(nms_pre=7, nms_post=7, nms_thr=0.9, score_thr=0.0), {}
    elif name == "border_targets":
        cfg, B, h, w, tc, kw = P, 2, 16, 24, dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0), dict(peaks=12)
    else:
        cfg = dataclasses.replace(synth.MUPOTS17, num_layers=3)
        B, h, w, tc, kw = 2, 24, 32, dict(nms_pre=9, nms_post=9, nms_thr=0.9, score_thr=0.0), {}
    case = util.make_case(cfg, B, h, w, seed=77, tc=tc, **kw)
    if name == "border_targets":
        for lv in case["levels"]:          # push the joint offsets so that targets and heads leave the map
            lv["pose_raw"][:, 3:3 + 3 * cfg.num_joints] *= 6.0
    ref, _ = util.run_oracle(case, tc)
    util.assert_margins(case, tc, ref)
    plan, got = util.run_gpu(case, tc, refine=True, refine_mode=1)
    compare(plan, got, ref)


def test_soft_oks_nms_matches_oracle():
    """test_cfg.nms_type != 'hard' -> soft_oks_nms (pose_nms.py:129-194): rescoring changes the order, nothing is dropped."""
    tc = dict(nms_pre=30, nms_post=12, nms_thr=0.9, score_thr=0.0, nms_type="soft")
    case = util.make_case(P, 3, 32, 48, seed=81, peaks=12, coherent=8, tc=tc)
    ref, _ = util.run_oracle(case, tc)
    util.assert_margins(case, tc, ref)
    plan, got = util.run_gpu(case, tc, refine=True)
    hard, _ = util.run_oracle(case, dict(tc, nms_type="hard"))
    assert any(a["index"].tolist() != b["index"].tolist() for a, b in zip(ref, hard)), "case does not exercise the rescoring"
    compare(plan, got, ref)
    for g in got:
        assert len(g["scores"]) == 12


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("refine", [True, False])
def test_fp16_head_outputs_are_read_natively(dtype, refine):
    """The reference's shipped fp16 mode (das_head.py:180,218 out_fp16=True, exp_panoptic.py:222) hands get_poses fp16
    maps.  The kernels read fp16 / bf16 cls, ctr and pose maps in place (das_levels.in_dtype) and compute in fp32: the
    result must equal the up-cast path bit for bit -- same people, same order, same coordinates."""
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 2, 24, 40, seed=91)
    dl = synth.levels_to(case["levels"], "cuda")
    head = DASHeadB200(num_joints=15, strides=[8], depth_factor=20, z_norm=50, root_idx=2, test_cfg=tc)
    head.load_refine_weights(synth.layers_to(case["layers"], "cuda"))
    low = lambda t: t.to(dtype)
    feats = [[f for f in lv["feats"]] for lv in dl]
    args_low = [[low(lv["cls"]) for lv in dl], [low(lv["pose_raw"]) for lv in dl], [low(lv["ctr"]) for lv in dl]]
    args_up = [[t.float() for t in a] for a in args_low]
    if refine:
        a = head.get_poses(*args_low, feats, case["metas"])
        b = head.get_poses(*args_up, feats, case["metas"])
    else:
        a = head.get_poses(*args_low, case["metas"])
        b = head.get_poses(*args_up, case["metas"])
    assert sum(len(x["scores"]) for x in a) > 0
    for x, y in zip(a, b):
        assert x["scores"] == y["scores"] and torch.equal(x["poses"], y["poses"]) and torch.equal(x["centers"], y["centers"])
        assert torch.equal(x["poses_cam"], y["poses_cam"])


def test_fp16_maps_through_the_host_entry():
    """das_plan_run_host with fp16 head-output maps in pinned host memory (element size 2 in the staging copies, in-place
    reads of the fp16 pose map, the centerness plane left on the host) == the device entry on the same fp16 maps, in all
    three transfer policies."""
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 3, 24, 40, seed=43, scales=(1.05, 0.95, 1.1, 0.9))
    plan = util.make_plan(case, tc)
    dl = synth.levels_to(case["levels"], "cuda")
    plan.bind([dict(cls=lv["cls"].half(), ctr=lv["ctr"].half(), pose=lv["pose_raw"].half(), feats=lv["feats"], scales=lv["scales"]) for lv in dl])
    plan.set_metas(case["metas"])
    plan.run()
    torch.cuda.synchronize()
    want = plan.output_block().clone()
    assert int(plan.views_of_block(want)["out_count"].sum()) > 0
    host_levels = [dict(cls=lv["cls"].half().pin_memory(), ctr=lv["ctr"].half().pin_memory(), pose=lv["pose_raw"].half().pin_memory(),
                        feats=[f.permute(0, 2, 3, 1).contiguous().pin_memory().permute(0, 3, 1, 2) for f in lv["feats"]],
                        scales=lv["scales"]) for lv in case["levels"]]
    for zero_copy, row_cache in ((False, False), (True, False), (True, True)):
        p2 = util.make_plan(case, tc)
        p2.set_host_mode(zero_copy, row_cache=row_cache)
        out = p2.alloc_host_out()
        for _ in range(2):
            p2.run_host(host_levels, case["metas"], out)
        torch.cuda.synchronize()
        a, b = plan.views_of_block(want), p2.views_of_block(p2.output_block().clone())
        for k in a:                                  # section by section: the 256-byte alignment gaps of a block are never written
            assert torch.equal(a[k], b[k]), (k, zero_copy, row_cache)
        res = p2.results(case["metas"], src=out)     # ... and what came back to the host arrays
        assert [len(r["scores"]) for r in res] == a["out_count"].tolist()


@pytest.mark.parametrize("variant", ["pyramid_pass_through", "k200_radix", "peak_mask", "odd_unaligned"])
def test_fp16_native_scan_paths(variant):
    """Every selection path of score_topk (bounded fast path, scratch keys + radix for K > 128, pass-through levels,
    3x3 peak mask, scalar loads for odd map sizes) on fp16 logit planes == the same planes up-cast to fp32."""
    kw = dict(pyramid_pass_through=dict(h=32, w=48, levels=4, tc=dict(nms_pre=60, nms_post=20, nms_thr=0.9, score_thr=0.0), peak=0),
              k200_radix=dict(h=40, w=56, levels=1, tc=dict(nms_pre=200, nms_post=50, nms_thr=0.9, score_thr=0.0), peak=0),
              peak_mask=dict(h=32, w=48, levels=1, tc=dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0), peak=3),
              odd_unaligned=dict(h=23, w=37, levels=1, tc=dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0), peak=0))[variant]
    tc = kw["tc"]
    strides = [8 * 2 ** l for l in range(kw["levels"])]
    case = util.make_case(dataclasses.replace(P, strides=tuple(strides)), 2, kw["h"], kw["w"], seed=93)
    dl = synth.levels_to(case["levels"], "cuda")
    assert len(dl) == kw["levels"]
    outs = []
    for up in (False, True):
        head = DASHeadB200(num_joints=15, strides=strides, depth_factor=20, z_norm=50, root_idx=2, test_cfg=tc,
                           peak_kernel=kw["peak"])
        conv = (lambda t: t.half().float()) if up else (lambda t: t.half())
        outs.append(head.get_poses([conv(lv["cls"]) for lv in dl], [conv(lv["pose_raw"]) for lv in dl], [conv(lv["ctr"]) for lv in dl],
                                   case["metas"]))
    assert sum(len(x["scores"]) for x in outs[0]) > 0
    for x, y in zip(*outs):
        assert x["scores"] == y["scores"] and torch.equal(x["poses"], y["poses"]) and torch.equal(x["centers"], y["centers"])


def test_caller_owned_output_block_and_fused_peer_stores():
    """das_plan_set_output_block: two plans write straight into slices of ONE staging buffer (no copy when results are
    collected); das_plan_set_peer_blocks: the NMS kernel also stores every result value into each peer's copy of the
    block -- here a second buffer on the same device stands in for a peer GPU -- and bumps the sequence word there."""
    tc = dict(nms_pre=10, nms_post=10, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(P, 3, 24, 40, seed=61, tc=tc)
    plan, got = util.run_gpu(case, tc, refine=True)
    want = plan.output_block().clone()
    nb, stride = want.numel(), plan.block_stride
    staging = torch.zeros((2, stride), dtype=torch.uint8, device="cuda")
    peer = torch.full((2, stride), 0xAB, dtype=torch.uint8, device="cuda")
    plans = []
    for k in range(2):
        p = util.make_plan(case, tc)
        dl = synth.levels_to(case["levels"], "cuda")
        p.bind([dict(cls=lv["cls"], ctr=lv["ctr"], pose=lv["pose_raw"], feats=lv["feats"], scales=lv["scales"]) for lv in dl])
        p.set_metas(case["metas"])
        p.set_output_block(staging[k])
        if k == 1:
            peer[1, nb:].zero_()
            p.set_peer_blocks([peer[1].data_ptr()])
        plans.append(p)
    for rep in range(3):                       # eager, capture, replay
        for p in plans:
            p.run()
    torch.cuda.synchronize()
    for k in range(2):
        assert torch.equal(staging[k, :nb], want), k
        assert plans[k].output_block().data_ptr() == staging[k].data_ptr()
        assert torch.equal(plans[k].t["out_pose"], plan.t["out_pose"])
    pv, wv = plan.views_of_block(peer[1, :nb]), plan.views_of_block(want)
    for name in wv:                                              # the "peer" received every result value ...
        assert torch.equal(pv[name], wv[name]), name
    assert int(peer[1, nb:nb + 4].view(torch.int32)) == 3        # ... and three sequence bumps
    assert int(staging[1, nb:nb + 4].view(torch.int32)) == 3
    assert torch.all(peer[0] == 0xAB)                            # nothing else was touched (nor the alignment gaps of peer[1])
    res = plans[1].results(case["metas"])
    for g, h in zip(got, res):
        assert g["scores"] == h["scores"] and torch.equal(g["poses_cam"], h["poses_cam"])
    # back to the plan's own block
    plans[0].set_output_block(None)
    plans[0].run()
    torch.cuda.synchronize()
    assert torch.equal(plans[0].output_block(), want) and plans[0].output_block().data_ptr() != staging[0].data_ptr()
