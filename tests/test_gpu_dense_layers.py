"""GPU: recursive_update.num_layers > 1 -- dense layers 1..L-1 followed by the sparse last layer,
against the oracle (which runs every layer densely like the reference) and the L=3 golden vectors."""
import os

import numpy as np
import pytest

import util
from das_b200 import synth
from oracle import make_golden as G
from test_gpu_parity import compare, TOL, GOLDEN

pytestmark = pytest.mark.gpu


# mode: refine_mode (None = tensor cores, 0 = fp32 SIMT); on_demand: layer L-2's sampling evaluated by the sparse last layer at
# the cells it looks at (the default on the tensor-core path) or over the whole map like the reference
@pytest.mark.parametrize("mode,on_demand", [(None, True), (None, False), (0, None)], ids=["tensor_core", "tensor_core_all_dense", "simt"])
@pytest.mark.parametrize("layers,J,root", [(2, 15, 2), (3, 17, 14), (2, 21, 14)])
def test_multi_layer_refinement_matches_oracle(layers, J, root, mode, on_demand):
    cfg = synth.HeadConfig(num_joints=J, root_idx=root, depth_factor=1.0, z_norm=50.0, num_layers=layers)
    tc = dict(nms_pre=20, nms_post=20, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(cfg, 2, 28, 36, seed=600 + layers, scales=(1.05, 0.95, 1.1, 0.9), tc=tc)
    ref, _ = util.run_oracle(case, tc)
    util.assert_margins(case, tc, ref)
    plan, got = util.run_gpu(case, tc, refine=True, refine_mode=mode, on_demand=on_demand)
    compare(plan, got, ref)


@pytest.mark.parametrize("on_demand", [True, False], ids=["on_demand", "all_dense"])
def test_multi_layer_pyramid(on_demand):
    cfg = synth.HeadConfig(num_joints=15, root_idx=2, depth_factor=20.0, z_norm=50.0, num_layers=2, strides=(8, 16, 32))
    tc = dict(nms_pre=50, nms_post=30, nms_thr=0.9, score_thr=0.03)
    case = util.make_case(cfg, 2, 32, 48, seed=700, peaks=20, tc=tc)
    ref, _ = util.run_oracle(case, tc)
    util.assert_margins(case, tc, ref)
    plan, got = util.run_gpu(case, tc, refine=True, on_demand=on_demand)
    compare(plan, got, ref)


@pytest.mark.parametrize("layers,force", [(2, 0), (3, 0), (3, 1), (3, 2)], ids=["L2", "L3", "L3_warp_per_item", "L3_batched"])
def test_on_demand_sampling_equals_the_dense_map(layers, force):
    """The sparse last layer evaluating layer L-2's sampling at its own cells (one device function shared with the dense
    kernel) returns the bits the whole-map pass returns -- both phase 1-2 kernels, eager and graph replay."""
    import torch
    from das_b200 import _lib
    cfg = synth.HeadConfig(num_joints=17, root_idx=14, depth_factor=1.0, z_norm=50.0, num_layers=layers)
    tc = dict(nms_pre=20, nms_post=20, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(cfg, 3, 40, 52, seed=640 + layers, peaks=12, tc=tc)
    lib = _lib.load()
    try:
        _lib.check(lib.das_debug_force_heads_kernel(force))
        outs = []
        for on in (True, False):
            plan, got = util.run_gpu(case, tc, refine=True, on_demand=on)
            plan.run(use_graph=True)         # second call: the captured graph
            torch.cuda.synchronize()
            outs.append(plan.results(case["metas"]))
    finally:
        lib.das_debug_force_heads_kernel(0)
    for a, b in zip(*outs):
        assert a.keys() == b.keys() and len(a["scores"]) > 0
        for k in a:
            assert torch.equal(a[k], b[k]) if torch.is_tensor(a[k]) else a[k] == b[k], k


@pytest.mark.parametrize("name", sorted(n for n in G.CASES if G.CASES[n][0].num_layers > 1))
def test_golden_vectors_with_dense_layers(name):
    """The reference's own outputs for the multi-layer cases: the synthetic J=17 / L=3 variant and the shipped MuPoTS model
    (J=21, L=2, 4-level pyramid, shipped test_cfg)."""
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg, levels, layers, metas, tc = G.build_case(name)
    case = dict(cfg=cfg, levels=levels, layers=layers, metas=metas, batch=levels[0]["cls"].shape[0])
    assert util.assert_margins(levels, tc) == int(gold["rank_margin_ulps"]) and float(gold["oks_margin"]) >= util.MIN_OKS_MARGIN
    plan, got = util.run_gpu(case, tc, refine=True)
    ci = plan.t["cand_index"].cpu()
    assert len(got) == int(gold["n_images"])
    for i, g in enumerate(got):
        lv, idx = util.slot_to_level_index(plan, g["slots"].cpu(), ci[i])
        assert idx == gold[f"index_{i}"].tolist() and lv == gold[f"level_{i}"].tolist()
        assert util.rel_err(g["poses"].cpu().numpy(), gold[f"poses_{i}"]) < TOL
        assert util.rel_err(g["poses_cam"].cpu().numpy(), gold[f"cam_{i}"]) < TOL


@pytest.mark.parametrize("layers", [2, 3])
def test_multi_layer_host_entry_equals_device_entry(layers):
    """das_plan_run_host with num_layers > 1: the dense layers' maps are bulk-copied in every transfer policy, the last layer
    reads its rows in place / through the row cache, and layer L-2's sampling is evaluated on demand there as well."""
    import torch
    cfg = synth.HeadConfig(num_joints=15, root_idx=2, depth_factor=20.0, z_norm=50.0, num_layers=layers)
    tc = dict(nms_pre=12, nms_post=12, nms_thr=0.9, score_thr=0.0)
    case = util.make_case(cfg, 2, 32, 40, seed=660 + layers, peaks=8, tc=tc)
    plan, _ = util.run_gpu(case, tc, refine=True)
    want = plan.views_of_block(plan.output_block().clone().cpu())
    assert int(want["out_count"].sum()) > 0
    host_levels = [dict(cls=lv["cls"].pin_memory(), ctr=lv["ctr"].pin_memory(), pose=lv["pose_raw"].pin_memory(),
                        feats=[f.permute(0, 2, 3, 1).contiguous().pin_memory().permute(0, 3, 1, 2) for f in lv["feats"]],
                        scales=lv["scales"]) for lv in case["levels"]]
    for zero_copy, row_cache in ((False, False), (True, False), (True, True)):
        p2 = util.make_plan(case, tc, refine=True)
        p2.set_host_mode(zero_copy, row_cache=row_cache)
        out = p2.alloc_host_out()
        for _ in range(3):                     # eager first run, then graph capture + replay
            p2.run_host(host_levels, case["metas"], out)
            for k, v in want.items():
                assert torch.equal(out[k], v), (zero_copy, row_cache, k)
