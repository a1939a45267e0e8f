"""Importing code by path.  Parity: `realhf/base/importing.py`.

`import_usercode` is how worker processes re-import the script that defined a custom experiment / interface / dataset
(`REAL_USER_CODE`, set by the quickstart entry point): registries are filled as a side effect of executing the module.
`import_package_modules` imports every module of a package directory whose file name matches a pattern (how a registry package
pulls in all of its implementations)."""

from __future__ import annotations

import importlib
import importlib.util
import os
import re
import sys
from types import ModuleType
from typing import List, Union


def import_usercode(path: str, module_name: str = "real_user_code") -> ModuleType:
    """Execute the python file at `path` as module `module_name`.  The module is put into `sys.modules` BEFORE it runs:
    dataclasses and pickling resolve classes through `sys.modules[cls.__module__]`."""
    path = os.path.abspath(path)
    if not os.path.isfile(path):
        raise FileNotFoundError(f"user code {path} does not exist")
    spec = importlib.util.spec_from_file_location(module_name, path)
    mod = importlib.util.module_from_spec(spec)
    prev = sys.modules.get(module_name)
    sys.modules[module_name] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        if prev is not None:
            sys.modules[module_name] = prev
        else:
            sys.modules.pop(module_name, None)
        raise
    return mod


def import_package_modules(package: str, pattern: Union[str, "re.Pattern"] = r"^(?!_).*\.py$") -> List[str]:
    """Import every module file of `package` (dotted name) whose file name matches `pattern`; returns the imported names."""
    pat = re.compile(pattern) if isinstance(pattern, str) else pattern
    pkg = importlib.import_module(package)
    done = []
    for d in getattr(pkg, "__path__", []):
        for fn in sorted(os.listdir(d)):
            if pat.match(fn) and fn.endswith(".py"):
                name = f"{package}.{fn[:-3]}"
                importlib.import_module(name)
                done.append(name)
    return done
