"""Every kernel family in one small pass with ragged shapes (tools/sanity_small.py): eval, streaming, feeder, training
with dropout + composite loss, FusedAdam, dilated training. compute-sanitizer is closed on this GPU pool, so this pass
plus the parity tests are the memory-safety net: TMA clips all tile tails, and a wrong offset shows up as a parity error."""
import os
import runpy

import pytest

pytestmark = pytest.mark.gpu


def test_small_end_to_end_pass(capsys):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, 'tools', 'sanity_small.py'), run_name='__main__')
    assert 'sanity pass done' in capsys.readouterr().out
