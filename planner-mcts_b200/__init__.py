"""B200-native rollout engine behind the BehaviorUCT* planner surface of
bark-simulator/planner-mcts.  Import as `planner_mcts_b200`
(the hyphenated directory name is not a Python identifier; the top-level
`planner_mcts_b200/` package re-points its __path__ here)."""
from . import _abi  # noqa: F401
from ._abi import BmctsError, default_config, load_library  # noqa: F401
from .engine import RESULT_DTYPE, RolloutEngine  # noqa: F401
