"""`qadence_extensions` — the plugin namespace qadence itself looks for
(`/root/reference/qadence/types.py:244-252`, `/root/reference/qadence/extensions.py:42-56,142-152`).

Shipping this package next to `qadence_b200` registers the B200-native backend under the name
"b200sv" with NO change to qadence:

* `qadence_extensions.types` extends `BackendName` / `Engine`;
* `qadence_extensions.extensions` supplies `available_backends / available_engines /
  supported_gates / set_backend_config`;
* `qadence_extensions.backends.b200sv.backend` is what `import_backend("b200sv")` falls back to;
* `import_config` and `import_engine` only look under `qadence.backends.<name>.config` and
  `qadence.engines.<engine>.differentiable_backend` (`extensions.py:30-39,71-80`), so a lazy
  meta-path alias maps those dotted names onto `qadence_b200.qadence_backend`.

**Replace mode.**  With `B200SV_REPLACE_PYQTORCH=1` in the environment the same alias mechanism also maps
`qadence.backends.pyqtorch.{backend,config}` and `qadence.engines.torch.differentiable_backend` onto the b200sv modules, and
the backend reports `name == "pyqtorch"`: every existing program — `QuantumModel(...)` with the default backend,
`backend_factory("pyqtorch", ...)` — then runs on the B200 kernels without a source change (and the reference's own test
suite can be pointed at this backend unmodified, tools/run_reference_suite.py).
"""

from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import sys

from .types import REPLACE_PYQTORCH

_ALIASES = {
    "qadence.backends.b200sv": "qadence_b200.qadence_backend",
    "qadence.backends.b200sv.backend": "qadence_b200.qadence_backend.backend",
    "qadence.backends.b200sv.config": "qadence_b200.qadence_backend.config",
    "qadence.engines.b200sv": "qadence_b200.qadence_backend",
    "qadence.engines.b200sv.differentiable_backend": "qadence_b200.qadence_backend.differentiable_backend",
}
if REPLACE_PYQTORCH:
    # only the three modules qadence resolves by name (extensions.py:30-80); `qadence.backends.pyqtorch` itself and its
    # `convert_ops` stay the reference's own (other reference modules import helpers from there)
    _ALIASES.update({
        "qadence.backends.pyqtorch.backend": "qadence_b200.qadence_backend.backend",
        "qadence.backends.pyqtorch.config": "qadence_b200.qadence_backend.config",
        "qadence.engines.torch.differentiable_backend": "qadence_b200.qadence_backend.differentiable_backend",
    })


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):  # type: ignore[no-untyped-def]
        if fullname in _ALIASES:
            return importlib.util.spec_from_loader(fullname, self, is_package=fullname.count(".") == 2)
        return None

    def create_module(self, spec):  # type: ignore[no-untyped-def]
        return importlib.import_module(_ALIASES[spec.name])

    def exec_module(self, module):  # type: ignore[no-untyped-def]
        return None


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    # FIRST on the meta path: once `qadence.backends.b200sv` is an alias of the real package, the stock PathFinder would
    # otherwise find `config.py` through the package's __path__ and execute it a second time under the alias name, giving
    # two distinct `Configuration` classes (backend_factory compares configuration types exactly, backends/api.py:34-41)
    sys.meta_path.insert(0, _AliasFinder())
