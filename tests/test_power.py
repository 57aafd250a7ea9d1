"""GpuPowerMonitor (hispmv_b200/power.py): the reference's power-monitor interface (common/src/fpga-power.cpp:9-57,
gpu/src/nvmlPower.cpp:51-91) driven by an injected reader on the CPU, and by NVML on the GPU box."""
import itertools
import time

import pytest

from hispmv_b200.power import GpuPowerMonitor, report


def _fake(values):
    it = itertools.cycle(values)
    return lambda device_id: (lambda: next(it))


def test_average_max_and_report_lines(tmp_path):
    mon = GpuPowerMonitor(period_s=0.02, reader_factory=_fake([100.0, 300.0]))
    assert mon.get_average_power() == (0.0, 0) and mon.get_max_power() == 0.0     # before any sample
    log = tmp_path / "power_logs" / "m.log"
    mon.start_monitoring(0, log_path=str(log))
    mon.start_monitoring(0)                                                         # second start is ignored
    time.sleep(0.3)
    mon.stop_monitoring()
    mon.stop_monitoring()
    avg, n = mon.get_average_power()
    assert n >= 4 and mon.get_max_power() == 300.0 and 100.0 <= avg <= 300.0
    lines = log.read_text().split()
    assert len(lines) == n and set(lines) <= {"100.000", "300.000"}                # watts, one per line
    text = report(mon).splitlines()
    assert text[0].startswith("Average Power: ") and text[0].endswith(" Watts")
    assert text[1] == "Max Power: 300 Watts" and text[2] == f"Number of Samples: {n}"
    n_after = mon.get_average_power()[1]
    time.sleep(0.1)
    assert mon.get_average_power()[1] == n_after                                    # stopped means stopped


def test_reader_errors_do_not_kill_the_sampler(capsys):
    calls = {"n": 0}

    def factory(device_id):
        def read():
            calls["n"] += 1
            if calls["n"] % 2:
                raise RuntimeError("nvml hiccup")
            return 42.0
        return read
    with GpuPowerMonitor(period_s=0.02, reader_factory=factory) as mon:
        mon.start_monitoring(0)
        time.sleep(0.25)
    assert mon.get_average_power()[0] == 42.0
    assert "Error retrieving power info" in capsys.readouterr().out


@pytest.mark.gpu
def test_nvml_reader_on_the_box():
    import torch
    mon = GpuPowerMonitor(period_s=0.05)
    mon.start_monitoring(0)
    a = torch.rand(8192, 8192, device="cuda")
    t0 = time.time()
    while time.time() - t0 < 0.5:
        a = a @ a
        a = a / a.max()
    torch.cuda.synchronize()
    mon.stop_monitoring()
    avg, n = mon.get_average_power()
    assert n >= 3 and 30.0 < avg < 1500.0 and mon.get_max_power() >= avg


def test_numa_binding_helper_never_raises():
    """sharded.bind_to_gpu_numa: without NVML (this container) or without a narrower GPU-local CPU set it reports
    'unchanged' and leaves the affinity mask alone."""
    import os
    from hispmv_b200.sharded import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    msg = bind_to_gpu_numa(0)
    assert isinstance(msg, str) and (msg.startswith("unchanged") or msg.startswith("bound to"))
    if msg.startswith("unchanged"):
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
