import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def _gpu_solver():
    """One GPU context for the whole session; fails loudly when libuwbgo.so or the GPU is missing."""
    from localization_b200 import Solver
    s = Solver(0)
    yield s
    s.close()


@pytest.fixture(params=["tile", "cta", "window"])
def solver(request, _gpu_solver):
    """Every parity test runs three times: through the tile kernels as the library picks them (lane = window,
    the large-batch path; 6x6 chains take the ITEM kernel of uwbgo_general_items.cu), through the tile kernels
    with the 6x6 CTA kernel forced for every chain (lm_general_cta_kernel: the kernel of forests, and the A/B
    partner of the ITEM kernel), and through the WINDOW path (one CTA per window; windows too large for it fall
    back to the tile kernels).
    `solver.path_ok(*tile_paths)` checks the path the last solve took in either mode."""
    s = _gpu_solver
    window = request.param == "window"
    s.set_window_path(2048 if window else 0)  # larger batches belong to the tile kernels in both modes
    s.path_ok = lambda *allowed: s.last_path in allowed or (window and s.last_path == 3)
    if request.param == "cta":
        os.environ["UWBGO_GENERAL_KERNEL"] = "cta"  # read by launch_solve at every launch
    yield s
    os.environ.pop("UWBGO_GENERAL_KERNEL", None)
    s.set_window_path(-1)
