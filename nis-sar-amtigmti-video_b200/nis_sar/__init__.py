"""nis_sar -- B200 (sm_100a) implementation of the SAR hot path of NIS-SAR-AMTIGMTI-Video:
raw-echo synthesis, Chirp Scaling focusing and two-channel DPCA/ATI, behind the reference's own
function names.  Importing this package does not touch CUDA; the compute entry points raise if the
CUDA library or a device is missing (no CPU fallback)."""
from .params import RadarParams, spaceborne_preset, airborne_vehicle_preset  # noqa: F401
from . import targets, scenes  # noqa: F401

_LAZY = {"run_bistatic_physics_gpu", "sar_focus_csa", "run_physics_engine", "run_moving_physics",
         "run_custom_physics", "gmti_products", "dpca_coregister", "install", "set_default_params",
         "save_ati_dpca_npz",
         "set_default_device"}


def __getattr__(name):
    if name in _LAZY:
        from . import api
        return getattr(api, name)
    if name in ("device", "api", "dist", "pipeline"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
