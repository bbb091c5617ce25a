"""Import helper: the package directory is named after the reference repository (hyphens, not importable as
an identifier), so it is registered under the module name `cmpc_b200`."""
import importlib.util
import os
import sys

PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                       "online-non-linear-centroidal-mpc-with-stability-guarantees-for-robust-locomotion-of-legged-robots-_b200")


def load():
    if "cmpc_b200" in sys.modules:
        return sys.modules["cmpc_b200"]
    spec = importlib.util.spec_from_file_location("cmpc_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["cmpc_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
