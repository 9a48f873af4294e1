"""Import alias: `import hgb200` == the package directory `single-person-pose-estimation_b200/`
(whose name is not a Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("single-person-pose-estimation_b200")
sys.modules[__name__] = _pkg
