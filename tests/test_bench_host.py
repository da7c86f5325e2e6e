"""Host-only checks of bench.py helpers (no GPU, no oracle)."""
import importlib.util
import io
import os
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_numa_binding_parses_sysfs_and_is_best_effort(monkeypatch):
    b = _bench()

    class Props:
        pci_domain_id, pci_bus_id, pci_device_id = 0, 0x1B, 0

    fake_torch = types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=lambda i: Props))
    allowed = sorted(os.sched_getaffinity(0))
    want = allowed[:2] if len(allowed) > 1 else allowed
    files = {
        "/sys/bus/pci/devices/0000:1b:00.0/numa_node": "1\n",
        "/sys/devices/system/node/node1/cpulist": ",".join(str(c) for c in want) + ",100000-100003\n",
    }
    real_open = open
    monkeypatch.setattr(b, "open", lambda p, *a, **k: io.StringIO(files[p]) if p in files else real_open(p, *a, **k),
                        raising=False)
    try:
        got = b.bind_to_gpu_numa_node(fake_torch, 0)
        assert got == {"node": 1, "cpus": len(want)}
        assert sorted(os.sched_getaffinity(0)) == want       # CPUs outside the allowed set are ignored
    finally:
        os.sched_setaffinity(0, allowed)
    # no NUMA information (node -1), unreadable sysfs, no CUDA device: stay unbound, never raise
    files["/sys/bus/pci/devices/0000:1b:00.0/numa_node"] = "-1\n"
    assert b.bind_to_gpu_numa_node(fake_torch, 0) is None
    del files["/sys/bus/pci/devices/0000:1b:00.0/numa_node"]
    assert b.bind_to_gpu_numa_node(fake_torch, 0) is None
    broken = types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=lambda i: 1 / 0))
    assert b.bind_to_gpu_numa_node(broken, 0) is None
    assert sorted(os.sched_getaffinity(0)) == allowed
