"""CPU: host-side logic of the drop-in head -- sharding, meta packing, output-block layout, and that
the product path refuses to run without its CUDA library/device (no silent fallback)."""
import os

import numpy as np
import pytest
import torch

from das_b200 import dist as ddist
from das_b200 import head as H
from das_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition_the_batch():
    for total in (1, 7, 64, 128):
        for world in (1, 2, 3, 4, 8):
            spans = [ddist.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_metas_layout():
    metas = synth.make_metas(3, 16, 24)
    sxy, cam = H.DecodePlan.pack_metas(metas)
    assert sxy.shape == (3, 2) and cam.shape == (3, 18) and sxy.dtype == np.float32 and cam.dtype == np.float64
    for b, m in enumerate(metas):
        assert np.array_equal(sxy[b], m["scale_factor"][:2])
        assert np.array_equal(cam[b, :3], m["cam"]["K"][0]) and np.array_equal(cam[b, 3:6], m["cam"]["K"][1])
        assert np.array_equal(cam[b, 6:15], m["cam"]["R"].reshape(9)) and np.array_equal(cam[b, 15:], m["cam"]["t"].reshape(3))
    # MuPoTS-style 2x3 intrinsics and a missing 'cam' (identity) are accepted
    sxy, cam = H.DecodePlan.pack_metas([dict(scale_factor=np.ones(4, np.float32), cam=dict(K=np.arange(6.).reshape(2, 3))),
                                        dict(scale_factor=np.ones(4, np.float32))])
    assert cam[0, :6].tolist() == [0, 1, 2, 3, 4, 5] and cam[1, 6:15].tolist() == np.eye(3).reshape(9).tolist()


def test_block_layout_is_aligned_and_roundtrips():
    B, P, J = 3, 10, 15
    lay, total = H.block_layout(B, P, J)
    assert total % 256 == 0 and all(off % 256 == 0 for *_, off in lay)
    block = torch.zeros(total, dtype=torch.uint8)
    v = H.block_views(block, B, P, J)
    v["out_count"][:] = torch.tensor([1, 2, 3], dtype=torch.int32)
    v["out_cam"][2, 9, 14, 2] = 7.5
    w = H.block_views(block.clone(), B, P, J)
    assert w["out_count"].tolist() == [1, 2, 3] and float(w["out_cam"][2, 9, 14, 2]) == 7.5
    assert w["out_pose"].shape == (B, P, J, 3) and w["out_world"].dtype == torch.float64


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda():
    head = H.DASHeadB200(num_joints=15, strides=(8,), depth_factor=20, z_norm=50, root_idx=2,
                         test_cfg=dict(nms_pre=10, nms_post=10))
    cls = torch.zeros(1, 1, 8, 8)
    pose = torch.zeros(1, 93, 8, 8)
    metas = synth.make_metas(1, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        head.get_poses([cls], [pose], [cls], metas)


def test_head_rejects_unsupported_config_keys():
    with pytest.raises(AssertionError):
        H.DASHeadB200(num_classes=2, num_joints=15, root_idx=2)
    with pytest.raises(AssertionError):
        H.DASHeadB200(num_joints=15, root_idx=2, recursive_update=dict(dim=2))


def test_product_package_never_imports_the_oracle():
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "das_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"


def test_sampler_order_and_interleave_match_multi_gpu_test():
    for size in (1, 5, 8, 13):
        for world in (1, 2, 4):
            parts = [[dict(i=i) for i in ddist.sampler_indices(size, world, r)] for r in range(world)]
            assert all(len(p) == len(parts[0]) for p in parts)
            merged = ddist.interleave_results(parts, size)
            assert [m["i"] for m in merged] == list(range(size))


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference`: exactly one JSON line on stdout with the contract's keys (the reference algorithm on
    the host cores; on a multi-rank launch only rank 0 prints)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, check=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoded_images_per_sec" and d["unit"] == "images/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1
    # the reference's own extracted functions when the tree or the oracle/_ref bundle is present, else the port
    from oracle import ref_extract as R
    assert d["cpu_baseline"]["kind"] == ("reference" if R.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # a non-zero rank of a torchrun launch exits without output
    env["RANK"] = "1"
    silent = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                            capture_output=True, text=True, timeout=600, env=env, check=True).stdout
    assert silent.strip() == ""


def test_get_poses_argument_forms():
    """The call forms get_poses accepts, resolved by content: the reference's (img_metas[, cfg[, rescale]]), positional or
    by keyword (das_head.py:653-659; detectors/das.py:37-38 splats the head outputs), and the extended
    (refine_feats, img_metas[, cfg[, rescale]]) one."""
    from das_b200.head import DASHeadB200
    metas = [dict(scale_factor=np.ones(4, np.float32), filename="a.jpg")]
    feats = [[object()]]
    sr = DASHeadB200._split_rest
    assert sr((metas,), None) == (None, metas, None)
    assert sr((metas, dict(nms_pre=5)), None) == (None, metas, dict(nms_pre=5))        # positional cfg
    assert sr((metas, dict(nms_pre=5), True), None) == (None, metas, dict(nms_pre=5))  # positional cfg, rescale
    assert sr((metas, None, True), dict(nms_pre=7)) == (None, metas, dict(nms_pre=7))
    assert sr((feats, metas), None) == (feats, metas, None)                            # extended call
    assert sr((feats, []), None) == (feats, [], None)                                  # empty batch
    assert sr((feats, metas, dict(nms_pre=5)), None) == (feats, metas, dict(nms_pre=5))          # extended + positional cfg
    assert sr((feats, metas, dict(nms_pre=5), True), None) == (feats, metas, dict(nms_pre=5))    # ... and rescale
    assert sr((feats, metas, None, False), dict(nms_pre=9)) == (feats, metas, dict(nms_pre=9))
    assert sr(([], dict(nms_pre=5)), None) == (None, [], dict(nms_pre=5))               # reference form, empty batch
    assert sr((), None, img_metas=metas) == (None, metas, None)                        # img_metas= by keyword
    assert sr((feats,), dict(nms_pre=3), img_metas=metas) == (feats, metas, dict(nms_pre=3))
    for bad in ((), (feats,), (feats, 3), (metas, None, None, None), (object(), object())):
        with pytest.raises(TypeError):
            sr(bad, None)
    with pytest.raises(TypeError):
        sr((feats, metas), None, img_metas=metas)


def test_pack_metas_fast_and_general_paths_agree():
    from das_b200 import synth
    from das_b200.head import DecodePlan
    metas = synth.make_metas(5, 16, 20, stride=8, seed=3)
    s1, c1 = DecodePlan.pack_metas(metas)
    partial = [dict(m) for m in metas]
    partial[2] = dict(scale_factor=metas[2]["scale_factor"], filename="x")            # no camera -> identity K, R; t = 0
    s2, c2 = DecodePlan.pack_metas(partial)
    assert np.array_equal(s1, s2) and s1.dtype == np.float32 and s1.flags["C_CONTIGUOUS"] and s1.shape == (5, 2)
    assert np.array_equal(np.delete(c1, 2, 0), np.delete(c2, 2, 0))
    assert c2[2].tolist() == [1, 0, 0, 0, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]
    two_by_three = [dict(m, cam=dict(K=m["cam"]["K"][:2], R=m["cam"]["R"], t=m["cam"]["t"])) for m in metas]   # MuPoTS K is 2x3
    assert np.array_equal(DecodePlan.pack_metas(two_by_three)[1], c1)


def test_head_maps_keep_half_precision_and_upcast_mixed_inputs():
    """fp16 / bf16 head outputs go to the kernels as they are (das_levels.in_dtype, read natively); anything mixed or
    exotic is up-cast to fp32 like mmcv's force_fp32 would (das_head.py:180,218 is the reference's fp16 mode)."""
    mk = lambda dt: [torch.zeros(2, 1, 4, 6, dtype=dt)]
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        c, r, p = H._head_maps(mk(dt), mk(dt), [torch.zeros(2, 93, 4, 6, dtype=dt)])
        assert c[0].dtype == r[0].dtype == p[0].dtype == dt
    c, r, p = H._head_maps(mk(torch.float16), mk(torch.float32), [torch.zeros(2, 93, 4, 6, dtype=torch.float16)])
    assert c[0].dtype == r[0].dtype == p[0].dtype == torch.float32
    c, r, p = H._head_maps(mk(torch.float64), mk(torch.float64), [torch.zeros(2, 93, 4, 6, dtype=torch.float64)])
    assert c[0].dtype == torch.float32
    from das_b200 import _lib
    assert (_lib.DTYPE_F32, _lib.DTYPE_F16, _lib.DTYPE_BF16) == (0, 1, 2)          # include/das_decode.h: DAS_DTYPE_*
    import re, os
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "das_decode.h")).read()
    for name, val in (("DAS_DTYPE_F32", 0), ("DAS_DTYPE_F16", 1), ("DAS_DTYPE_BF16", 2)):
        assert re.search(r"#define\s+%s\s+%d\b" % (name, val), hdr)


def test_bench_rank_pinning_gives_disjoint_core_slices():
    """bench.pin_rank_to_cores: every rank of a node gets its own slice of the allowed cores (and the call never raises)."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    before = os.sched_getaffinity(0)
    try:
        world = min(4, max(1, len(before)))
        slices = []
        for r in range(world):
            os.sched_setaffinity(0, before)
            info = bench.pin_rank_to_cores(r, world)
            assert "error" not in info, info
            slices.append(frozenset(os.sched_getaffinity(0)))
        assert all(s and s <= before for s in slices)
        if len(before) >= world:
            assert all(a.isdisjoint(b) for i, a in enumerate(slices) for b in slices[i + 1:])
    finally:
        os.sched_setaffinity(0, before)


def test_reference_citations_point_inside_the_cited_files():
    """Parity is argued through `file.py:line` citations of the reference; none may point past the end of its file."""
    import importlib.util
    ref = os.environ.get("DAS_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("reference tree not mounted (GPU box)")
    spec = importlib.util.spec_from_file_location("check_citations", os.path.join(ROOT, "tools", "check_citations.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n, bad = mod.check(ref)
    assert n > 150 and not bad, bad
