"""Import alias: the package directory is named `zk-state-proofs_b200` (not a valid Python
identifier), so `import zk_state_proofs_b200` resolves to it through this shim."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "zk-state-proofs_b200")]
__spec__.submodule_search_locations = __path__  # makes this module a package for relative imports
__package__ = __name__
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _f.name, "exec"))
