"""Registry-shaped front end of the B200 CI plugins.

Mirrors the part of ``qdk_chemistry.algorithms`` (python/src/qdk_chemistry/algorithms/
registry.py:295-372 create, :496-540 register, :543-590 available, :593-632 show_default,
:635-663 unregister) that a user of the ``macis_cas`` / ``macis_asci`` / ``macis_pmc``
calculators touches, on top of the compiled ``_core`` module (pybind11 over the C++ plugin
layer in host/). The calculators run on the GPU through libb2ci.so; nothing here computes.
"""
from __future__ import annotations

from typing import Callable

from ._core import algorithms as _alg
from ._core.algorithms import (  # noqa: F401  (re-exported)
    B200Asci,
    B200Cas,
    B200Pmc,
    DuplicateRegistrationError,
    MultiConfigurationCalculator,
    MultiConfigurationCalculatorFactory,
    ProjectedMultiConfigurationCalculator,
    ProjectedMultiConfigurationCalculatorFactory,
    clear_communicator,
    davidson_solver,
    last_run_stats,
    row_block,
    set_communicator,
    set_device,
)

_FACTORIES = {
    "multi_configuration_calculator": MultiConfigurationCalculatorFactory,
    "projected_multi_configuration_calculator": ProjectedMultiConfigurationCalculatorFactory,
}
# names a QDK user already has in scripts -> the drop-in that replaces them
_DROP_IN = {"macis_cas": "b200_cas", "macis_asci": "b200_asci", "macis_pmc": "b200_pmc"}


def _factory(algorithm_type: str):
    try:
        return _FACTORIES[algorithm_type]
    except KeyError:
        raise KeyError(f"Algorithm type '{algorithm_type}' is not available; this build provides "
                       f"{sorted(_FACTORIES)}") from None


def create(algorithm_type: str, algorithm_name: str | None = None, **kwargs):
    """Create an algorithm instance; keyword arguments are forwarded to ``settings().update``.
    The reference names ``macis_cas`` / ``macis_asci`` / ``macis_pmc`` resolve to their B200
    drop-ins unless an algorithm of that exact name has been registered."""
    f = _factory(algorithm_type)
    name = algorithm_name or ""
    if name in _DROP_IN and not f.has(name):
        name = _DROP_IN[name]
    algo = f.create(name)
    if kwargs:
        algo.settings().update(kwargs)
    return algo


def available(algorithm_type: str | None = None):
    if algorithm_type is None:
        return {t: sorted(f.available()) for t, f in _FACTORIES.items()}
    return sorted(_factory(algorithm_type).available())


def show_default(algorithm_type: str | None = None):
    if algorithm_type is None:
        return {t: f.default_algorithm_name() for t, f in _FACTORIES.items()}
    return _factory(algorithm_type).default_algorithm_name()


def register(generator: Callable[[], object]) -> None:
    """Register a generator of MultiConfigurationCalculator instances (Python subclasses go
    through the trampoline, exactly as in the reference)."""
    probe = generator()
    _factory(probe.type_name()).register_instance(generator)


def unregister(algorithm_type: str, algorithm_name: str) -> None:
    if not _factory(algorithm_type).unregister_instance(algorithm_name):
        raise KeyError(f"Algorithm '{algorithm_name}' of type '{algorithm_type}' is not registered")


def inspect_settings(algorithm_type: str, algorithm_name: str):
    s = create(algorithm_type, algorithm_name).settings()
    return [(k, s.get_type_name(k), s.get(k), s.get_description(k) if s.has_description(k) else None)
            for k in s.keys()]


def init_distributed_from_torch(device: int | None = None) -> None:
    """One process per GPU: take rank / world size from an initialised ``torch.distributed``
    process group, create the NCCL unique id on rank 0, broadcast it, and hand it to the plugin
    layer. Afterwards every ``run`` shards determinant rows over the ranks."""
    import torch
    import torch.distributed as dist

    from . import device as _dev
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.cuda.current_device()
    set_device(device)
    if world == 1:
        clear_communicator()
        return
    uid = broadcast_unique_id(_dev.Context.comm_unique_id() if rank == 0 else None,
                              f"cuda:{device}" if dist.get_backend() == "nccl" else "cpu")
    set_communicator(uid, rank, world)


def broadcast_unique_id(unique_id: bytes | None, tensor_device: str = "cpu") -> bytes:
    """Rank 0 passes the 128-byte NCCL unique id, the others None; returns it on every rank."""
    import torch
    import torch.distributed as dist
    idt = torch.zeros(128, dtype=torch.uint8, device=tensor_device)
    if dist.get_rank() == 0:
        if unique_id is None or len(unique_id) != 128:
            raise ValueError("rank 0 must pass the 128-byte unique id")
        idt.copy_(torch.frombuffer(bytearray(unique_id), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().numpy().tobytes())
