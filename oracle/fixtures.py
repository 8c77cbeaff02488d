"""Seeded synthetic inputs live in the repo-level module `synthetic` (numpy only): they are the WORKLOAD of
the benchmark and the tests, not part of the oracle.  This alias keeps `from oracle import fixtures` working for the tests and
the golden-vector scripts."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthetic import *            # noqa: F401,F403,E402
from synthetic import _rng         # noqa: F401,E402
