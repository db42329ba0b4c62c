"""ORACLE — test infrastructure only.

CPU restatement of the reference's hot path (emNavi/AirGym, airgym/envs/base/hovering.py et al.).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (airgym_b200/) never does and fails loudly without its CUDA library.
"""
from .spec import QuadSpec, CTL_MODES, TASKS  # noqa: F401
from .hovering import HoveringOracle, TrackingOracle, make_oracle  # noqa: F401
