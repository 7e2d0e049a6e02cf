"""of_dis_b200/numa.py: placement helper for the pinned staging buffers (pure sysfs + affinity)."""
import os

from of_dis_b200 import numa


def test_cpulist_parser():
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()
    assert numa._parse_cpulist("5") == {5}


def test_bind_is_a_no_op_without_topology_and_always_restorable():
    before = os.sched_getaffinity(0)
    node, prev = numa.bind_to_gpu_node(0)  # no GPU / no sysfs entry here: nothing changes
    try:
        assert prev is None or set(prev) == set(before)
        if node is None:
            assert os.sched_getaffinity(0) == before
    finally:
        numa.unbind(prev)
    assert os.sched_getaffinity(0) == before
