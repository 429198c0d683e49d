"""nis_sar -- B200 (sm_100a) implementation of the SAR hot path of NIS-SAR-AMTIGMTI-Video:
raw-echo synthesis, Chirp Scaling focusing and two-channel DPCA/ATI, behind the reference's own
function names.  Importing this package does not touch CUDA; the compute entry points raise if the
CUDA library or a device is missing (no CPU fallback)."""
from .params import RadarParams, spaceborne_preset, airborne_vehicle_preset  # noqa: F401
from . import targets, scenes  # noqa: F401

_SUBMODULES = ("device", "api", "dist", "hostio", "viewer", "video")


def __getattr__(name):
    """Everything public in ``nis_sar.api`` (the reference's entry points: run_bistatic_physics_gpu, sar_focus_csa,
    sar_focus_rda, gmti_products, add_ocean_noise, tdbp_gpu, install, ...) is reachable as ``nis_sar.<name>``; torch and the
    CUDA library are only imported when one of them is first used."""
    if name in _SUBMODULES:
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    if not name.startswith("_"):
        from . import api
        if hasattr(api, name):
            return getattr(api, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
