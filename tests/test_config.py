"""network.json of the reference (src/config.rs:5-9) and its mapping onto one multi-GPU box"""
import json
import os

import pytest

from distributed_plonk_b200.config import NetworkConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_local8_sample():
    cfg = NetworkConfig.load(os.path.join(ROOT, "config", "network.local8.json"))
    assert cfg.n_workers == 8 and cfg.workers_on("127.0.0.1") == list(range(8))
    assert cfg.gpu_plan("127.0.0.1", 8) == {i: i for i in range(8)}
    env = cfg.rendezvous_env(3)
    assert env == {"RANK": "3", "WORLD_SIZE": "8", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "8900"}
    with pytest.raises(ValueError):
        cfg.gpu_plan("127.0.0.1", 4)
    with pytest.raises(ValueError):
        cfg.rendezvous_env(8)


def test_reference_shaped_file_and_rejects(tmp_path):
    p = tmp_path / "network.json"      # two workers on two boxes, the shape of the reference's own file
    p.write_text(json.dumps({"slaves": ["10.0.0.201:8888", "10.0.0.202:9999"], "peers": ["10.0.0.201:8899", "[::1]:9988"]}))
    cfg = NetworkConfig.load(str(p))
    assert cfg.slaves == [("10.0.0.201", 8888), ("10.0.0.202", 9999)] and cfg.peers[1] == ("::1", 9988)
    assert cfg.workers_on("10.0.0.202") == [1] and cfg.gpu_plan("10.0.0.202", 8) == {1: 0}
    for bad in ({"slaves": ["10.0.0.1:1"], "peers": []}, {"slaves": ["host.example:80"], "peers": ["10.0.0.1:1"]},
                {"slaves": ["10.0.0.1"], "peers": ["10.0.0.1:1"]}, {"slaves": ["10.0.0.1:70000"], "peers": ["10.0.0.1:1"]}):
        p.write_text(json.dumps(bad))
        with pytest.raises(ValueError):
            NetworkConfig.load(str(p))


def test_gpu_numa_affinity_helper(tmp_path):
    """parallel.near_gpu: the bench's staging buffers are allocated on the CPUs next to the GPU (sysfs cpulist)
    and the previous affinity comes back afterwards"""
    import os

    from distributed_plonk_b200 import parallel

    assert parallel.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert parallel.parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    one = {min(before)}
    with parallel.near_gpu(0, cpus=one) as what:
        inside = os.sched_getaffinity(0)
    assert os.sched_getaffinity(0) == before
    if len(before) > 1:
        assert inside == one and "NUMA" in what
    with parallel.near_gpu(0, cpus=set(before)) as what:          # nothing to narrow: left alone
        assert os.sched_getaffinity(0) == before and what.startswith("unchanged")
    with parallel.near_gpu(0, cpus={1 << 20}) as what:            # a list that shares no CPU with ours: left alone
        assert os.sched_getaffinity(0) == before and what.startswith("unchanged")
    with parallel.near_gpu(0) as what:                            # no GPU here: unknown topology, left alone
        assert os.sched_getaffinity(0) == before
