"""Hooks qadence imports from `qadence_extensions.extensions` (qadence/extensions.py:142-152)."""

from __future__ import annotations

from typing import Any


def _q() -> Any:
    import qadence.extensions as qe

    return qe


def available_backends() -> dict:
    """Installed backends; those whose third-party simulator is missing are skipped, not fatal."""
    qe = _q()
    from qadence.types import BackendName

    out = {}
    for name in BackendName.list():
        try:
            out[BackendName(name)] = qe.import_backend(name)
        except Exception:
            continue
    return out


def available_engines() -> dict:
    qe = _q()
    from qadence.types import Engine

    out = {}
    for name in Engine.list():
        try:
            out[Engine(name)] = qe.import_engine(name)
        except Exception:
            continue
    return out


def supported_gates(backend_name: str) -> list:
    return _q()._supported_gates(backend_name)


def set_backend_config(backend: Any, diff_mode: Any) -> None:
    """b200sv always uses gate-level (uuid) parameters; every other backend keeps qadence's rule."""
    qe = _q()
    if str(backend.name) == "b200sv":
        qe._validate_diff_mode(backend, diff_mode)
        backend.config._use_gate_params = True
        return
    qe._set_backend_config(backend, diff_mode)
