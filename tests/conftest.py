import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """One HotPathEngine for the whole GPU session (fails loudly if the CUDA library is missing)."""
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from slaf_b200.engine import HotPathEngine

    eng = HotPathEngine(device=0, max_rows=1 << 23, max_cells=1 << 17)
    yield eng
    eng.close()


def pytest_sessionfinish(session, exitstatus):
    """With a checked build of the library (SLAFB_LIB=build_checked/libslaf_b200_checked.so, scripts/checked_build.sh) report what
    the device-side bounds checks saw during the whole session; a violation fails the run."""
    if not os.environ.get("SLAFB_LIB"):
        return
    try:
        import ctypes

        from slaf_b200 import _native

        info = (ctypes.c_uint32 * 8)()
        n = _native.lib().slafb_debug_check_failures(info)
    except Exception as err:  # pragma: no cover
        print(f"\n[checked build] could not read the check counters: {err!r}")
        return
    if n == -1:
        return
    line = f"[checked build] device-side bounds checks violated: {n} (first: code {info[1]}, value {info[2]}, CTA {info[3]}, thread {info[4]})"
    print("\n" + line)
    out = os.environ.get("SLAFB_CHECK_REPORT")
    if out:
        with open(out, "a") as f:
            f.write(line + f"; pytest exit status {int(exitstatus)}\n")
    if n != 0:
        session.exitstatus = 1
