"""CPU: host logic, C-ABI surface, URDF compiler, error behaviour, multi-rank sharding over gloo."""
import os
import re
import socket

import numpy as np
import pytest

import helpers
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hrp_b200 import arch, consts, synth, urdf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "hrp_b200.h")).read()
    declared = set(re.findall(r"\b(hrp_[a-z0-9_]+)\s*\(", hdr)) - {"hrp_status"}
    from hrp_b200 import capi
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib_built, name), name
    assert b"sm_100a" in lib_built.hrp_version()


@pytest.mark.parametrize("robot", ["panda", "kuka", "baxter"])
@pytest.mark.parametrize("backbone", ["resnet50", "hrnet32"])
def test_cxx_graph_builder_wants_the_reference_state_dict(lib_built, robot, backbone):
    from hrp_b200.model import HoliRobPoseB200
    m = HoliRobPoseB200(robot, {"backbone_name": backbone})
    mine = sorted((n, tuple(s)) for n, s, _ in arch.full_net(robot, backbone))
    assert sorted(m.expected_tensors()) == mine


@pytest.mark.parametrize("name", list(helpers.VARIANT_CASES))
def test_cxx_graph_builder_wants_the_variant_state_dicts(lib_built, name):
    """Constructor variants (8f N4): same tensor list as the reference built with those switches (the goldens' generator
    loads synth's state dict into it with strict=True)."""
    from hrp_b200.model import HoliRobPoseB200
    cfg, ctor = helpers.VARIANT_CASES[name]
    m = HoliRobPoseB200("panda", dict(cfg))
    mine = sorted((n, tuple(s)) for n, s, _ in arch.full_net("panda", "resnet50", {k: v for k, v in ctor.items() if k not in ("depth_root", "joint_bounds")}))
    assert sorted(m.expected_tensors()) == mine


def test_unsupported_configs_fail_loudly(lib_built):
    from hrp_b200.model import HoliRobPoseB200
    with pytest.raises(ValueError):
        HoliRobPoseB200("owi535")
    with pytest.raises(NotImplementedError):
        HoliRobPoseB200("panda", {"use_rpmg": True})
    with pytest.raises(NotImplementedError):                                  # full_net.py:377 reads the ResNet trunk's map
        HoliRobPoseB200("panda", {"reg_joint_map": True, "joint_conv_dim": [64, 64, 64], "backbone_name": "hrnet32"})
    with pytest.raises(NotImplementedError):
        HoliRobPoseB200("panda", {"reg_joint_map": True, "joint_conv_dim": [100, 64, 64]})
    with pytest.raises(ValueError, match="not in list"):                      # kps_need_depth.index(reference_keypoint_id), full_net.py:328
        HoliRobPoseB200("panda", {"multi_kp": True, "kps_need_depth": [0, 1]})
    with pytest.raises(NotImplementedError):
        HoliRobPoseB200("panda", {"rotation_dim": 4})
    with pytest.raises(NotImplementedError):
        HoliRobPoseB200("panda", {"backbone_name": "resnet34"})
    m = HoliRobPoseB200("panda")
    with pytest.raises(RuntimeError, match="unexpected"):
        m.load_state_dict({"bogus.weight": np.zeros(3, np.float32)})
    m2 = HoliRobPoseB200("panda")
    with pytest.raises(RuntimeError):
        m2.forward_record(torch.zeros(1), torch.zeros(1), torch.zeros(1), torch.zeros(1))


def test_set_weight_rejects_bad_shapes(lib_built):
    import ctypes as C
    from hrp_b200 import capi
    from hrp_b200.model import HoliRobPoseB200
    m = HoliRobPoseB200("kuka")
    a = np.zeros((64, 3, 3, 3), np.float32)      # reg_backbone.conv1 is 7x7 for resnet50
    shape = (C.c_int64 * 4)(*a.shape)
    st = lib_built.hrp_set_weight(m._h, b"reg_backbone.conv1.weight", a.ctypes.data_as(C.c_void_p), shape, 4, 0)
    assert st == -4 and b"expected 7" in lib_built.hrp_last_error()
    st = lib_built.hrp_set_weight(m._h, b"nope", a.ctypes.data_as(C.c_void_p), shape, 4, 0)
    assert st == -4
    assert lib_built.hrp_finalize_weights(m._h) == -4     # tensors missing
    assert lib_built.hrp_forward(m._h, 0, 0, 0, 0, 1, 0, 0) == -3   # not finalized


def test_synthetic_state_dict_is_deterministic_and_complete():
    a = synth.make_state_dict("panda", "resnet50")
    b = synth.make_state_dict("panda", "resnet50")
    assert list(a) == [n for n, _, _ in arch.full_net("panda", "resnet50")]
    for k in ("reg_backbone.layer3.2.conv2.weight", "rootnet_backbone.stage4.1.fuse_layers.2.0.1.0.weight",
              "rootnet_backbone.stage3.0.branches.1.2.bn1.running_var", "fc_rot_1.weight"):
        assert np.array_equal(a[k], b[k])
    assert a["rootnet_backbone.bn1.running_var"].min() > 0
    np.testing.assert_allclose(a["init_pose"][0], consts.ROBOTS["panda"]["init_pose"], rtol=1e-7)
    np.testing.assert_array_equal(a["init_rot"][0], consts.INIT_ROT6D)


@pytest.mark.parametrize("robot", ["panda", "kuka", "baxter"])
def test_urdf_program(robot):
    R, P = urdf.load_robot(robot)
    spec = consts.ROBOTS[robot]
    assert [j.name for j in R.actuated] == spec["joints"]
    assert P.dof == spec["dof"] and P.nkpt == spec["nkpt"] and P.n_slots == 0
    assert sorted(P.kp_index) == list(range(P.nkpt))
    assert P.kp_step == sorted(P.kp_step)
    # only joints on a keypoint path survive: Panda's fingers and Baxter's head are pruned
    assert len(P.step_type) == {"panda": 7, "kuka": 7, "baxter": 14}[robot]


BRANCHY = """<robot name="tree">
<link name="b"/><link name="l1"/><link name="l2"/><link name="l3"/><link name="r2"/><link name="r3"/><link name="tip"/>
<joint name="j1" type="revolute"><parent link="b"/><child link="l1"/><origin xyz="0 0 0.2" rpy="0.1 0 0"/><axis xyz="0 0 1"/></joint>
<joint name="j2" type="revolute"><parent link="l1"/><child link="l2"/><origin xyz="0.1 0 0" rpy="0 0.3 0"/><axis xyz="0 1 0"/></joint>
<joint name="j3" type="prismatic"><parent link="l2"/><child link="l3"/><origin xyz="0 0.1 0"/><axis xyz="1 0 0"/></joint>
<joint name="k2" type="revolute"><parent link="l1"/><child link="r2"/><origin xyz="-0.1 0 0" rpy="0 0 0.5"/><axis xyz="1 0 0"/></joint>
<joint name="k3" type="revolute"><parent link="r2"/><child link="r3"/><origin xyz="0 0 0.3"/><axis xyz="0 0.6 0.8"/></joint>
<joint name="f" type="fixed"><parent link="r3"/><child link="tip"/><origin xyz="0 0.05 0.02" rpy="0.2 0.1 0"/></joint>
</robot>"""


def test_urdf_branching_tree_uses_saved_frames():
    R = urdf.Robot(BRANCHY)
    frames = [("b", np.zeros(3)), ("l3", np.array([0.0, 0.0, 0.01])), ("r2", np.zeros(3)), ("tip", np.zeros(3))]
    P = urdf.compile_program(R, frames, 2)
    assert P.n_slots == 1 and urdf.PARENT_BASE in P.step_parent and 0 in P.step_parent
    assert [j.name for j in R.actuated] == ["j1", "j2", "k2", "j3", "k3"]


def test_shard_range_covers_batch():
    from hrp_b200 import dist as hd
    for total in (1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            spans = [hd.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hrp_b200 import dist as hd
    B, dof, nkpt = 3, 8, 7
    offs = hd.record_offsets(B, dof, nkpt)
    rec = torch.zeros(offs[-1])
    for f, w in enumerate(hd.field_widths(dof, nkpt)):
        rec[offs[f]:offs[f] + B * w] = torch.arange(B * w, dtype=torch.float32) + 1000 * f + 100000 * rank
    out = hd.gather_records(rec, B, dof, nkpt)
    ok = out["joint_angles"].shape == (world * B, dof) and out["kp2d_fk"].shape == (world * B, nkpt, 2)
    for r in range(world):
        ok = ok and float(out["trans"][r * B, 0]) == 2000 + 100000 * r
        ok = ok and float(out["kp3d_fk"][r * B + 1, 0, 0]) == 7000 + nkpt * 3 + 100000 * r
    # the multi_kp variant's record carries one more field (every regressed depth)
    offs = hd.record_offsets(B, dof, nkpt, depth_num=3)
    rec = torch.zeros(offs[-1])
    for f, w in enumerate(hd.field_widths(dof, nkpt, 3)):
        rec[offs[f]:offs[f] + B * w] = torch.arange(B * w, dtype=torch.float32) + 1000 * f + 100000 * rank
    out = hd.gather_records(rec, B, dof, nkpt, depth_num=3)
    ok = ok and out["depths"].shape == (world * B, 3) and float(out["depths"][B, 1]) == 10001 + 100000 * 1
    ok = ok and float(out["kp2d_fk"][B, 0, 0]) == 9000 + 100000 * 1
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_records_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_generated_fk_chains_are_current():
    """csrc/fk_programs_gen.h must be exactly what scripts/gen_fk_programs.py emits from the packaged URDFs."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_fk_programs.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_committed_launch_list_reproduces_the_committed_traffic_summary(tmp_path):
    """profiles/r02_dram_traffic_per_step.json (read by bench.py for roofline.traffic) is what scripts/ncu_launches.py
    makes of the committed ncu launch list -- the evidence chain stays reproducible without a GPU."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, "profiles", "r02_launches_f16_b64_one_forward.csv")
    committed = json.load(open(os.path.join(root, "profiles", "r02_dram_traffic_per_step.json")))
    out = tmp_path / "t.json"
    subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_launches.py"), csv_path, str(committed["launches"]), str(out)],
                   check=True, capture_output=True)
    again = json.load(open(out))
    assert again["launches"] == committed["launches"]
    assert abs(again["dram_bytes_per_step"] - committed["dram_bytes_per_step"]) < 1.0
    assert abs(again["l2_bytes_per_step"] - committed["l2_bytes_per_step"]) < 1.0
    conv = sum(v for k, v in again["by_kernel_dram_mb"].items() if k.startswith("conv_"))
    assert conv > 0.9 * again["dram_bytes_per_step"] / 1e6        # the conv family is what moves the bytes
