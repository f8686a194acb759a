"""B200-native configuration-interaction hot path for QDK/Chemistry.

``device``      thin wrappers over the C ABI (libb2ci.so)
``algorithms``  the MultiConfigurationCalculator plugin surface (needs the compiled ``_core``)
``data``        Hamiltonian / Settings / Configuration / Wavefunction stand-ins
``workloads``   seeded synthetic integrals for the BASELINE configs

Sub-modules are imported lazily so that CPU-only tooling can import the package without the
compiled extensions; using them without the extensions raises (there is no fallback).
"""
import importlib

__all__ = ["device", "algorithms", "data", "workloads"]


def __getattr__(name):
    if name in __all__:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
