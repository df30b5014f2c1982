"""Problem builders shared by the tests: they live in the package (sse_b200.problems) so that
bench.py does not depend on the test tree; this module re-exports them."""
from sse_b200.problems import *  # noqa: F401,F403
from sse_b200.problems import rough_state  # noqa: F401
