"""Import shim: the package lives in `planner-mcts_b200/` (hyphen, as the task's
layout names it); this module makes it importable as `planner_mcts_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "planner-mcts_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
