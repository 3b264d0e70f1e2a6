"""Stage sections: an NVTX range per pipeline stage (always, for profilers) and opt-in stage timers (ST_TIMING=1 or
timing.enable()): synchronising wall-clock sections used by tools/stage_stats.py and bench.py's stage table.
Timers disabled = no synchronisation."""
import os
import time
from contextlib import contextmanager

import torch

ENABLED = bool(int(os.environ.get("ST_TIMING", "0")))
RECORDS = {}
SAMPLES = {}


def enable(on=True):
    global ENABLED
    ENABLED = on
    RECORDS.clear()
    SAMPLES.clear()


NVTX = bool(int(os.environ.get("ST_NVTX", "1")))      # NVTX ranges around every pipeline stage (profilers; ~1 us each)


@contextmanager
def section(name):
    nvtx = NVTX and torch.cuda.is_available()
    if nvtx:
        torch.cuda.nvtx.range_push(name)
    try:
        if not ENABLED:
            yield
            return
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        yield
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        RECORDS[name] = RECORDS.get(name, 0.0) + dt
        SAMPLES.setdefault(name, []).append(dt)
    finally:
        if nvtx:
            torch.cuda.nvtx.range_pop()
