"""The driver's smoke() entry point must stay green: it is run on a fresh B200 before the bench."""
import pytest


@pytest.mark.gpu
def test_graft_entry_smoke():
    import __graft_entry__ as g
    g.smoke()
