"""CPU: the host I/O pipeline of the eval adapters keeps order, overlaps, and surfaces failures."""
import threading
import time

import pytest

from lerf_pytorch_b200.eval_common import IoPipeline


@pytest.mark.parametrize("threads", [0, 1, 4])
def test_prefetch_keeps_order_and_submit_returns_results(threads):
    io = IoPipeline(threads)
    seen = []

    def load(i):
        time.sleep(0.002 * ((7 * i) % 5))  # finish out of order
        return i * i

    futs = []
    for v in io.prefetch(load, range(23)):
        seen.append(v)
        futs.append(io.submit(lambda x: x + 1, v))
    io.drain()
    assert seen == [i * i for i in range(23)]
    assert [f.result() for f in futs] == [i * i + 1 for i in range(23)]
    io.close()


def test_prefetch_runs_ahead_but_bounded():
    io = IoPipeline(threads=3, depth=2)
    started, lock = [], threading.Lock()

    def load(i):
        with lock:
            started.append(i)
        return i

    it = io.prefetch(load, range(10))
    assert next(it) == 0
    time.sleep(0.05)
    with lock:
        ahead = max(started)
    assert 1 <= ahead <= 2  # items 1..depth are being decoded while item 0 is consumed, nothing beyond
    assert list(it) == list(range(1, 10))
    io.close()


def test_failures_surface_in_drain():
    io = IoPipeline(threads=2)

    def boom():
        raise OSError("disk full")

    io.submit(boom)
    with pytest.raises(OSError):
        io.drain()
    io.close()
