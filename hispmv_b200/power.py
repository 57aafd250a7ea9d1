"""Board power sampler for the GPU engine, with the reference's monitor interface.

Mirrors FpgaPowerMonitor (common/include/fpga-power.h:17-40, common/src/fpga-power.cpp:9-57: startMonitoring /
stopMonitoring / getAveragePower / getMaxPower, one sample per second from a background thread) and the V100
benchmark's NVML poller (gpu/src/nvmlPower.cpp:51-91: nvmlDeviceGetPowerUsage in milliwatts, one value in watts per
line of ./power_logs/<matrix>.log).  Samples come from NVML (the `pynvml` module of nvidia-ml-py); a different reader
can be injected, which is how the CPU tests drive it.
"""
from __future__ import annotations

import os
import threading
import time
from typing import Callable, List, Optional, Tuple


def nvml_reader(device_id: int) -> Callable[[], float]:
    """A callable returning the board's present draw in watts (nvmlDeviceGetPowerUsage / 1000)."""
    import pynvml
    pynvml.nvmlInit()
    handle = pynvml.nvmlDeviceGetHandleByIndex(device_id)

    def read() -> float:
        return pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0
    return read


class GpuPowerMonitor:
    def __init__(self, period_s: float = 1.0, reader_factory: Callable[[int], Callable[[], float]] = nvml_reader):
        self.period_s = period_s
        self._factory = reader_factory
        self._samples: List[float] = []
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._log = None
        self.debug = False

    # -- FpgaPowerMonitor's interface -----------------------------------------------------------------
    def start_monitoring(self, device_id: int, debug: bool = False, log_path: Optional[str] = None) -> None:
        if self._thread is not None:
            return                                              # fpga-power.cpp:10: a second start is ignored
        read = self._factory(device_id)
        self.debug = debug
        if log_path:
            os.makedirs(os.path.dirname(os.path.abspath(log_path)), exist_ok=True)
            self._log = open(log_path, "w")
        self._stop.clear()
        self._thread = threading.Thread(target=self._poll, args=(read,), daemon=True)
        self._thread.start()

    def stop_monitoring(self) -> None:
        if self._thread is None:
            return
        self._stop.set()
        self._thread.join()
        self._thread = None
        if self._log:
            self._log.close()
            self._log = None

    def get_average_power(self) -> Tuple[float, int]:
        """(mean watts, number of samples); (0.0, 0) before the first sample (fpga-power.cpp:24-30)."""
        with self._lock:
            n = len(self._samples)
            return (sum(self._samples) / n if n else 0.0), n

    def get_max_power(self) -> float:
        with self._lock:
            return max(self._samples) if self._samples else 0.0

    def samples(self) -> List[float]:
        with self._lock:
            return list(self._samples)

    # -- sampling thread: half a period, read, half a period (fpga-power.cpp:41-52, nvmlPower.cpp:66-88) -----
    def _poll(self, read: Callable[[], float]) -> None:
        while not self._stop.is_set():
            if self._stop.wait(self.period_s / 2):
                break
            try:
                w = float(read())
            except Exception as ex:  # noqa: BLE001
                print(f"Error retrieving power info: {ex}")
                w = None
            if w is not None:
                with self._lock:
                    self._samples.append(w)
                if self.debug:
                    print(f"sample: {w}")
                if self._log:
                    self._log.write(f"{w:.3f}\n")
            self._stop.wait(self.period_s / 2)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop_monitoring()


def report(monitor: GpuPowerMonitor) -> str:
    """The three lines the reference prints after a power run (common/src/spmv-helper.cpp:1041-1049)."""
    avg, n = monitor.get_average_power()
    return f"Average Power: {avg:g} Watts\nMax Power: {monitor.get_max_power():g} Watts\nNumber of Samples: {n}"


def energy_per_run_mj(avg_watts: float, run_us: float) -> float:
    """Energy of one SpMV in millijoules: the efficiency column of the reference's tables (GFLOPS/W = gflops / avg W)."""
    return avg_watts * run_us * 1e-3
