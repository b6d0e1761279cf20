"""Import alias: `import ssb200` == the package in simple-spectral_b200/ (whose directory name is not a
valid Python identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("simple-spectral_b200")
