"""The C-ABI library loads on a machine without a GPU, exports every symbol include/sai_b200_osc.h
declares, fails loudly instead of falling back to the CPU, and its built-in models agree with the
independently written oracle tables."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle.robots import make_chain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "sai_b200_osc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(osc_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from sai_primitives_b200 import capi
    declared = _declared_functions()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
        assert name in capi.SYMBOLS, "capi.py does not bind %s" % name
    assert sorted(capi.SYMBOLS) == declared
    assert lib.osc_abi_version() == capi.OSC_ABI_VERSION


def test_struct_layouts_match_header_sizes(lib):
    """defaults round-trip through the C structs: a layout mismatch would scramble them"""
    from sai_primitives_b200 import capi
    p = capi.MftParams()
    assert lib.osc_mft_default_params(C.byref(p)) == 0
    assert list(p.kp_pos) == [100.0] * 3 and list(p.kv_ori) == [28.3] * 3       # MotionForceTask.h:44-49
    assert (p.kp_force, p.kv_force, p.ki_force, p.kff_force) == (0.7, 10.0, 1.3, 0.95)
    assert (p.s_min, p.s_max, p.bie_threshold) == (6e-3, 6e-2, 0.1)
    assert p.dynamic_decoupling_type == capi.BOUNDED_INERTIA_ESTIMATES and p.buffer_size == 200
    assert p.singularity_handling_enabled == 1 and p.passivity_enabled == 0
    j = capi.JointParams()
    assert lib.osc_joint_default_params(C.byref(j)) == 0
    assert list(j.kp) == [50.0] * 8 and list(j.kv) == [14.0] * 8 and j.bie_threshold == 0.1   # JointTask.h:32-37
    assert j.dynamic_decoupling_type == capi.BOUNDED_INERTIA_ESTIMATES


@pytest.mark.parametrize("name", ["panda", "rrrr", "puma_like", "panda_sliding_base"])
def test_builtin_models_match_oracle_tables(lib, name):
    from sai_primitives_b200 import capi
    d = capi.ModelDesc()
    assert lib.osc_builtin_model(name.encode(), C.byref(d)) == 0
    ch = make_chain(name)
    n = ch.n
    assert d.n == n
    assert list(d.jtype[:n]) == list(ch.jtype)
    for i in range(n):
        assert np.allclose(np.array(d.axis[i][:]), ch.axis[i], atol=0)
        assert np.allclose(np.array(d.R_fix[i][:]).reshape(3, 3), ch.R_fix[i], atol=1e-15)
        assert np.allclose(np.array(d.t_fix[i][:]), ch.t_fix[i], atol=1e-15)
        assert abs(d.mass[i] - ch.mass[i]) < 1e-15
        assert np.allclose(np.array(d.com[i][:]), ch.com[i], atol=1e-15)
        assert np.allclose(np.array(d.inertia[i][:]).reshape(3, 3), ch.inertia[i], atol=1e-15)
    assert np.allclose(d.q_lower[:n], ch.q_lower) and np.allclose(d.q_upper[:n], ch.q_upper)
    assert np.allclose(d.effort[:n], ch.effort) and np.allclose(d.dq_max[:n], ch.dq_max)
    for link, (body, R, t) in ch.link_frames.items():
        f = capi.LinkFrame()
        assert lib.osc_builtin_link(name.encode(), link.encode(), C.byref(f)) == 0
        assert f.body == body
        assert np.allclose(np.array(f.R[:]).reshape(3, 3), R, atol=1e-15) and np.allclose(np.array(f.t[:]), t, atol=1e-15)
    assert lib.osc_builtin_model(b"no_such_robot", C.byref(d)) != 0


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly (there is no CPU path behind the ABI)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the no-device failure mode cannot be observed here")
    import sai_primitives_b200 as sp
    with pytest.raises(sp.capi.OscError) as e:
        sp.BatchedRobot("panda", 4)
    assert e.value.code == sp.capi.OSC_ERR_NO_DEVICE
    assert "no CPU fallback" in e.value.message


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under sai_primitives_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sai_primitives_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboscref" not in txt and "osc_ref.cpp" not in txt, f


def test_shard_range_c_and_python_agree(lib):
    """osc_shard_range (C callers) and sharding.shard_range (Python callers): the same contiguous partition, every robot owned once"""
    import ctypes as C
    from sai_primitives_b200.sharding import shard_of, shard_range
    for n_total in (0, 1, 7, 8, 65536, 65537, 262144, 1000003):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                first, count = C.c_int64(-1), C.c_int64(-1)
                assert lib.osc_shard_range(n_total, r, world, C.byref(first), C.byref(count)) == 0
                lo, hi = shard_range(n_total, r, world)
                assert (first.value, first.value + count.value) == (lo, hi)
                assert lo == covered
                covered = hi
                if hi > lo:
                    assert shard_of(lo, n_total, world) == r and shard_of(hi - 1, n_total, world) == r
            assert covered == n_total
    first, count = C.c_int64(), C.c_int64()
    assert lib.osc_shard_range(10, 4, 4, C.byref(first), C.byref(count)) != 0
