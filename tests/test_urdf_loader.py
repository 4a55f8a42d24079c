"""the URDF loader of the library (host code, no GPU needed).  A URDF text rendered from the oracle's
robot descriptions must give the same model as the library's built-in table of that robot, and -- where the reference
tree is present (this container, not the GPU box) -- the reference's own URDF data files must give the built-in models."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import robots as OR


def _urdf_text(desc):
    out = ['<?xml version="1.0" ?>', '<!-- rendered by tests/test_urdf_loader.py -->', '<robot name="%s">' % desc["name"]]
    links = desc["links"]
    for l in links:
        I = np.asarray(l["inertia"])
        out.append('  <link name="%s">\n    <inertial>\n      <origin xyz="%r %r %r" rpy="0 0 0"/>\n      <mass value="%r"/>' % ((l["name"],) + tuple(l["com"]) + (l["mass"],)))
        out.append('      <inertia ixx="%r" iyy="%r" izz="%r" ixy="%r" ixz="%r" iyz="%r"/>\n    </inertial>' % tuple(float(v) for v in (I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2])))
        out.append('    <visual><origin xyz="1 2 3"/><geometry><box size="1 1 1"/></geometry><material name="m"><color rgba="0 0 0 1"/></material></visual>\n  </link>')
    out.insert(3, '  <link name="world_anchor"/>')          # a massless root, like the reference files that start with a base link
    prev = "world_anchor"
    for k, l in enumerate(links):
        jt = l["jtype"]
        out.append('  <joint name="joint%d" type="%s">\n    <parent link="%s"/>\n    <child link="%s"/>' % (k, jt, prev, l["name"]))
        out.append('    <origin xyz="%r %r %r" rpy="%r %r %r"/>' % (tuple(l["xyz"]) + tuple(l["rpy"])))
        if jt != "fixed":
            lo, hi, vel, eff = l["limits"]
            out.append('    <axis xyz="%r %r %r"/>\n    <limit lower="%r" upper="%r" velocity="%r" effort="%r"/>' % (tuple(l["axis"]) + (lo, hi, vel, eff)))
        out.append("  </joint>")
        prev = l["name"]
    out.append("</robot>")
    return "\n".join(out)


def _desc_arrays(lib, capi, name):
    d = capi.ModelDesc()
    assert lib.osc_builtin_model(name.encode(), C.byref(d)) == 0
    n = d.n
    f = lambda a, *shape: np.array(a[:]).reshape(-1, *shape)[:n] if shape else np.array(a[:n])
    return dict(n=n, jtype=np.array(d.jtype[:n]), axis=np.array([list(r) for r in d.axis])[:n], R_fix=np.array([list(r) for r in d.R_fix])[:n],
                t_fix=np.array([list(r) for r in d.t_fix])[:n], mass=np.array(d.mass[:n]), com=np.array([list(r) for r in d.com])[:n],
                inertia=np.array([list(r) for r in d.inertia])[:n], q_lower=np.array(d.q_lower[:n]), q_upper=np.array(d.q_upper[:n]),
                dq_max=np.array(d.dq_max[:n]), effort=np.array(d.effort[:n]))


def _same_model(a, b, tol=1e-12):
    assert a["n"] == b["n"]
    for k in a:
        if k != "n":
            assert np.abs(np.asarray(a[k], dtype=float) - np.asarray(b[k], dtype=float)).max() <= tol, k


@pytest.mark.parametrize("name", ["panda", "panda_sliding_base", "rrrr", "puma_like"])
def test_urdf_text_gives_the_builtin_model(name):
    from sai_primitives_b200 import capi
    import sai_primitives_b200 as sp
    lib = capi.load_library()
    desc = OR.DESCRIPTIONS[name]()
    sp.registerUrdf("urdf_" + name, _urdf_text(desc))
    _same_model(_desc_arrays(lib, capi, "urdf_" + name), _desc_arrays(lib, capi, name))
    for l in desc["links"]:                       # link frames (a link behind fixed joints hangs off its parent body)
        fa, fb = capi.LinkFrame(), capi.LinkFrame()
        assert lib.osc_builtin_link(("urdf_" + name).encode(), l["name"].encode(), C.byref(fa)) == 0
        assert lib.osc_builtin_link(name.encode(), l["name"].encode(), C.byref(fb)) == 0
        assert fa.body == fb.body and np.allclose(fa.R[:], fb.R[:], atol=1e-12) and np.allclose(fa.t[:], fb.t[:], atol=1e-12)


def test_urdf_errors():
    import sai_primitives_b200 as sp
    with pytest.raises(ValueError, match="serial"):
        sp.registerUrdf("tree", '<robot name="t"><link name="a"/><link name="b"/><link name="c"/>'
                                '<joint name="j1" type="revolute"><parent link="a"/><child link="b"/></joint>'
                                '<joint name="j2" type="revolute"><parent link="a"/><child link="c"/></joint></robot>')
    with pytest.raises(ValueError, match="unknown link"):
        sp.registerUrdf("bad", '<robot name="t"><link name="a"/><joint name="j" type="fixed"><parent link="a"/><child link="zz"/></joint></robot>')
    with pytest.raises(ValueError, match="cannot open"):
        sp.registerUrdf("nofile", "/nonexistent/robot.urdf")


REF = "/root/reference/examples"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,path", [("panda", "15-haptic_control_impedance_type/panda_arm.urdf"),
                                       ("rrrr", "11-planar_robot_controller/rrrrbot.urdf"),
                                       ("panda_sliding_base", "06-partial_joint_task/panda_arm_sliding_base.urdf")])
def test_reference_urdf_files_give_the_builtin_models(name, path):
    """the built-in tables were typed in from these files (csrc/builtin_models.cpp header): read the files themselves"""
    from sai_primitives_b200 import capi
    import sai_primitives_b200 as sp
    lib = capi.load_library()
    sp.registerUrdf("ref_" + name, os.path.join(REF, path))
    _same_model(_desc_arrays(lib, capi, "ref_" + name), _desc_arrays(lib, capi, name), tol=1e-12)


def test_world_file_and_path_prefixes(tmp_path):
    """SaiModel::URDF_FOLDERS / ReplaceUrdfPathPrefix and the robot's base pose from the world file
    (examples/01-joint_control/01-joint_control.cpp:40-71, examples/15-...cpp:124-126)"""
    import ctypes as C
    from sai_primitives_b200 import capi
    lib = capi.load_library()
    from oracle.robots import rpy_to_matrix, rrrr_description
    (tmp_path / "models").mkdir()
    (tmp_path / "models" / "arm.urdf").write_text(_urdf_text(rrrr_description()))
    (tmp_path / "world.urdf").write_text('''<?xml version="1.0" ?>
<world name="demo_world" gravity="0.0 0.0 -1.62">
  <robot name="OTHER"><model dir="${MODELS}" path="missing.urdf" name="x" /><origin xyz="9 9 9" rpy="0 0 0" /></robot>
  <robot name="ARM">
    <model dir="${MODELS}" path="arm.urdf" name="rrrr" />
    <origin xyz="0.1 -0.2 0.3" rpy="0.2 -0.4 1.1" />
  </robot>
  <static_object name="Floor"><origin xyz="0 0 0" rpy="0 0 0" /></static_object>
</world>''')
    assert lib.osc_urdf_set_folder(b"MODELS", str(tmp_path / "models").encode()) == 0
    assert lib.osc_urdf_set_folder(b"WORLD_DIR", str(tmp_path).encode()) == 0
    buf = C.create_string_buffer(512)
    assert lib.osc_urdf_replace_path_prefix(b"${MODELS}/arm.urdf", buf, 512) == 0
    assert buf.value.decode() == str(tmp_path / "models" / "arm.urdf")
    assert lib.osc_urdf_replace_path_prefix(b"${NOT_SET}/arm.urdf", buf, 512) != 0
    g = (C.c_double * 3)()
    assert lib.osc_world_register_robot(b"${WORLD_DIR}/world.urdf", b"ARM", b"world_arm", g) == 0, lib.osc_urdf_last_error()
    assert list(g) == [0.0, 0.0, -1.62]
    d = capi.ModelDesc()
    assert lib.osc_builtin_model(b"world_arm", C.byref(d)) == 0
    assert d.n == 4
    assert np.abs(np.array(d.R_world_base[:]).reshape(3, 3) - rpy_to_matrix((0.2, -0.4, 1.1))).max() < 1e-15
    assert list(d.t_world_base[:]) == [0.1, -0.2, 0.3]
    base = _desc_arrays(lib, capi, "rrrr"); got = _desc_arrays(lib, capi, "world_arm")
    for k in ("axis", "t_fix", "mass", "com", "inertia"):
        assert np.abs(np.asarray(base[k]) - np.asarray(got[k])).max() < 1e-12, k
    assert lib.osc_world_register_robot(b"${WORLD_DIR}/world.urdf", b"NOBODY", b"w2", None) != 0
    assert lib.osc_world_register_robot(b"${WORLD_DIR}/world.urdf", b"OTHER", b"w3", None) != 0      # its URDF does not exist
