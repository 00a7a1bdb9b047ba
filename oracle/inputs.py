"""TEST INFRASTRUCTURE - the seeded synthetic batches live in the package
(`dolfinx_external_operator_b200.synthetic`, NumPy only) so that bench.py's GPU arm does not
import anything from oracle/; re-exported here for the golden generator and the tests."""
from dolfinx_external_operator_b200.synthetic import *  # noqa: F401,F403
from dolfinx_external_operator_b200.synthetic import heat_batch, isihara_batch, mc_batch, mc_demo_path, vm_batch  # noqa: F401
