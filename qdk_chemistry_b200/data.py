"""Data stand-ins at the plugin boundary (see host/include/qdk_b200/data.hpp)."""
from ._core.data import (  # noqa: F401
    Configuration,
    Hamiltonian,
    Settings,
    SettingNotFound,
    SettingsAreLocked,
    SettingTypeMismatch,
    UserSettings,
    Wavefunction,
)
