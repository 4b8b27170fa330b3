"""C-ABI surface (loads without a GPU, exports everything include/*.h declares) and host-side logic."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch

from mycobotgym_b200 import _lib, flatten, vector_env

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "mycobot_b200.h")).read()
    declared = set(re.findall(r"\b(mcb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_the_header(lib):
    assert lib.mcb_model_desc_size() == ctypes.sizeof(flatten.ModelDesc)
    assert lib.mcb_task_cfg_size() == ctypes.sizeof(flatten.TaskCfg)
    lib.mcb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.mcb_version()


def test_errors_are_reported_not_thrown(lib):
    lib.mcb_last_error.restype = ctypes.c_char_p
    assert lib.mcb_model_create(None, 0, None) < 0
    assert b"null" in lib.mcb_last_error()
    assert lib.mcb_step(None, None, None, None, None, None, None, None, None, None, None) < 0
    assert lib.mcb_batch_num_envs(None) == -1


def test_registry_matches_reference_registration():
    reg = vector_env.registry()
    assert len(reg) == 30                                              # __init__.py:5-35: 36 combos - 6 fetch+joint
    kw = reg["MyCobotPickAndPlace-Sparse-joint-v0"]
    assert kw == dict(model_path="./assets/mycobot280.xml", reward_type="sparse", has_object=True, controller_type="joint",
                      fetch_env=False, max_episode_steps=50)
    assert reg["MyCobotReach-Dense-mocap-v0"]["model_path"] == "./assets/mycobot280_mocap.xml"
    assert "MyCobotFetchReach-Dense-joint-v0" not in reg
    with pytest.raises(KeyError):
        vector_env.make("NoSuchEnv-v0")
    with pytest.raises(NotImplementedError):
        vector_env.make("MyCobotReach-Dense-joint-v1")


def test_unbuilt_rows_fail_loudly():
    for kwargs in (dict(controller_type="mocap"), dict(model_path="./assets/mycobot280_mocap.xml"), dict(controller_type="osc")):
        with pytest.raises(ValueError):                      # controller and model variant must agree
            vector_env.MyCobotVectorEnv(num_envs=1, **kwargs)
    with pytest.raises(AssertionError):                      # mycobot.py:96
        vector_env.MyCobotVectorEnv(num_envs=1, fetch_env=True, controller_type="joint")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure path")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vector_env.MyCobotVectorEnv(num_envs=2)


def test_spaces_and_sharding():
    b = vector_env.Box(-1.0, 1.0, (7,), np.float32)
    b.seed(0)
    x = b.sample()
    assert x.dtype == np.float32 and b.contains(x)
    spans = [vector_env.shard_envs(131072, r, 8) for r in range(8)]
    assert spans[0] == (0, 16384) and spans[-1] == (114688, 16384)
    spans = [vector_env.shard_envs(10, r, 4) for r in range(4)]
    assert [s[1] for s in spans] == [3, 3, 2, 2] and spans[3][0] + spans[3][1] == 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    start, cnt = vector_env.shard_envs(1000, rank, world)
    stats = torch.tensor([cnt, rank, 1.5 * cnt, 50.0 * cnt, cnt * 10.0, 0, 0, 0], dtype=torch.float64)
    vector_env.all_reduce_stats(stats)
    q.put((rank, stats.tolist()))
    dist.destroy_process_group()


def test_stats_all_reduce_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res[0] == res[1]
    assert res[0][0] == 1000 and res[0][1] == 1 and res[0][2] == 1500.0 and res[0][4] == 10000.0
