"""One process (one torch import): tools/pf_diag.py at the C3 shape, then the prefilter tests."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import pf_diag

try:
    pf_diag.main()
except Exception as e:  # keep going: the tests say more
    print("pf_diag failed:", repr(e))
sys.stdout.flush()
import pytest

sys.exit(pytest.main(["-m", "gpu", "-q", "--timeout", "60", str(ROOT / "tests" / "test_search_prefilter.py")]))
