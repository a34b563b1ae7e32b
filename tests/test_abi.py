"""CPU: the C-ABI library loads without a GPU and exports every symbol include/das_decode.h declares;
struct layouts of the ctypes binding equal the C compiler's; argument validation fails loudly."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from das_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "das_decode.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(das_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in das_decode.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) <= set(names)
    assert b"sm_100a" in lib.das_version()


def test_struct_layouts_match_the_c_compiler():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "das_decode.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(das_level_desc), sizeof(das_levels), sizeof(das_decode_cfg),
         sizeof(das_buffers), offsetof(das_levels, lv), offsetof(das_level_desc, H), offsetof(das_decode_cfg, dataset_depth_factor));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [C.sizeof(_lib.LevelDesc), C.sizeof(_lib.Levels), C.sizeof(_lib.DecodeCfg), C.sizeof(_lib.Buffers),
            _lib.Levels.lv.offset, _lib.LevelDesc.H.offset, _lib.DecodeCfg.dataset_depth_factor.offset]
    assert got == want


def test_slot_rules_follow_the_reference_topk_condition():
    lib = _lib.load()
    # das_head.py:716-717: top-k only if nms_pre > 0 and N > nms_pre
    assert lib.das_level_slots(128, 208, 10) == 10
    assert lib.das_level_slots(2, 4, 10) == 8          # pass-through level
    assert lib.das_level_slots(2, 5, 10) == 10         # N == nms_pre: not ">" -> pass-through of all 10
    assert lib.das_level_slots(4, 4, -1) == 16
    lv = _lib.Levels()
    lv.n_levels, lv.batch = 3, 1
    for l, (h, w) in enumerate([(32, 48), (16, 24), (2, 3)]):
        lv.lv[l].H, lv.lv[l].W = h, w
    assert lib.das_candidate_slots(C.byref(lv), 100) == 100 + 100 + 6
    assert lib.das_output_slots(206, 30) == 30 and lib.das_output_slots(20, 100) == 20 and lib.das_output_slots(20, -1) == 20


def test_argument_errors_are_reported_not_swallowed():
    lib = _lib.load()
    cfg = _lib.DecodeCfg(num_joints=99, root_idx=0, num_heads=4, feat_channels=256, num_layers=1, nms_pre=10, nms_post=10)
    shape = _lib.Levels()
    shape.n_levels, shape.batch = 1, 1
    shape.lv[0].H, shape.lv[0].W, shape.lv[0].stride = 8, 8, 8
    plan = C.c_void_p()
    st = lib.das_plan_create(C.byref(cfg), C.byref(shape), C.byref(plan))
    assert st == -3 and b"num_joints" in lib.das_last_error()          # DAS_ERR_CAPACITY
    cfg.num_joints, cfg.root_idx = 15, 20
    assert lib.das_plan_create(C.byref(cfg), C.byref(shape), C.byref(plan)) == -1    # DAS_ERR_ARG
    cfg.root_idx, cfg.nms_pre = 2, 5000
    assert lib.das_plan_create(C.byref(cfg), C.byref(shape), C.byref(plan)) == -3
    cfg.nms_pre, cfg.refine, cfg.feat_channels = 10, 1, 200
    assert lib.das_plan_create(C.byref(cfg), C.byref(shape), C.byref(plan)) == -4    # DAS_ERR_UNSUPPORTED
    with pytest.raises(_lib.DasError):
        _lib.check(-1, "x")
    assert lib.das_score_topk(None, None, 10, 0, None, None, 10, None, None) == -1


def test_c_host_links_against_the_library_and_fails_loudly_without_a_device():
    """examples/decode_host.c: a C99 caller needs nothing but include/das_decode.h and the shared library; where no CUDA
    device exists the plan constructor reports DAS_ERR_CUDA with the failing runtime call (there is no CPU path)."""
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    _lib.load()
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "decode_host")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "examples", "decode_host.c"), "-L", lib_dir, "-ldas_decode",
                               "-Wl,-rpath," + lib_dir, "-Wl,--allow-shlib-undefined", "-o", exe])
        r = subprocess.run([exe], capture_output=True, text=True)
    assert "sm_100a" in r.stdout and "candidate slots per image: 10, output slots: 10" in r.stdout
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0 and "plan ready: 10 candidate slots, 10 output slots" in r.stdout, r.stderr
    else:
        assert r.returncode == 2 and "das_plan_create failed (-2)" in r.stderr and "cuda" in r.stderr.lower(), r.stderr
