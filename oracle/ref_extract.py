"""TEST INFRASTRUCTURE ONLY -- run the *unmodified* reference functions in this container.

The reference simulators are top-level scripts: importing one executes the whole
simulation (and needs matplotlib, which is absent).  So the functions on the hot
path are pulled out with ``ast`` -- only the ``FunctionDef`` nodes are compiled --
and executed in a namespace that supplies the module globals they read
(SURVEY.md section 8c).  No reference source is copied into this repository; the
files are read from ``/root/reference`` at call time, which exists only in the
build container.  Used by ``oracle/make_golden.py`` (fixture generation) and by
the optional ``tests/test_oracle_vs_reference.py`` (skipped when the reference
tree is absent, e.g. on the GPU box).
"""
from __future__ import annotations

import ast
import contextlib
import io
import os

import numpy as np

REFERENCE_DIR = os.environ.get("NIS_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "sar_ati_dcpa_sim_csa.py"))


def _extract(filename: str, names: tuple[str, ...], namespace: dict) -> dict:
    path = os.path.join(REFERENCE_DIR, filename)
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in wanted}
    if missing:
        raise RuntimeError(f"{filename}: functions not found: {sorted(missing)}")
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    return {n: namespace[n] for n in names}


def _quiet(fn):
    """The reference prints progress lines; swallow them."""
    def wrapped(*a, **k):
        with contextlib.redirect_stdout(io.StringIO()):
            return fn(*a, **k)
    return wrapped


def ati_csa_functions(g: dict):
    """``run_bistatic_physics_gpu`` and ``sar_focus_csa`` of sar_ati_dcpa_sim_csa.py
    (:106-181, :202-396) bound to the globals in ``g`` (keys C, R0, FC, BW, T_p, FS)."""
    import torch
    ns = {"np": np, "torch": torch, "device": torch.device("cpu")}
    ns.update({k: g[k] for k in ("C", "R0", "FC", "BW", "T_p", "FS")})
    f = _extract("sar_ati_dcpa_sim_csa.py", ("run_bistatic_physics_gpu", "sar_focus_csa"), ns)
    return _quiet(f["run_bistatic_physics_gpu"]), _quiet(f["sar_focus_csa"])


def satellite_engine(g: dict):
    """``run_physics_engine`` of sar_satellite_sim.py:211-305 (globals C, R0, FC, BW, T_p)."""
    ns = {"np": np}
    ns.update({k: g[k] for k in ("C", "R0", "FC", "BW", "T_p")})
    return _quiet(_extract("sar_satellite_sim.py", ("run_physics_engine",), ns)["run_physics_engine"])


def moving_engine(g: dict):
    """``run_moving_physics`` of sar_satellite_moving_sim.py:111-159."""
    ns = {"np": np}
    ns.update({k: g[k] for k in ("C", "R0", "FC", "BW", "T_p")})
    return _quiet(_extract("sar_satellite_moving_sim.py", ("run_moving_physics",), ns)["run_moving_physics"])


def vehicle_engine(g: dict):
    """``run_custom_physics`` of sar_vehicle_sim.py:83-126 (globals C, R0)."""
    ns = {"np": np}
    ns.update({k: g[k] for k in ("C", "R0")})
    return _quiet(_extract("sar_vehicle_sim.py", ("run_custom_physics",), ns)["run_custom_physics"])


def ati_inline_products(slc1, slc2):
    """Execute the reference's *inline* GMTI statements (sar_ati_dcpa_sim_csa.py:414-419 and
    :447-449 -- they are module-level code, not a function) on the given SLC pair and return
    the names they bind."""
    path = os.path.join(REFERENCE_DIR, "sar_ati_dcpa_sim_csa.py")
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = {414, 415, 416, 418, 419, 447, 448, 449}
    body = [n for n in tree.body if getattr(n, "lineno", -1) in keep]
    targets = set()
    for n in body:
        for t in getattr(n, "targets", []):
            targets.add(getattr(t, "id", None) or getattr(getattr(t, "value", None), "id", None))
    expect = {"ati_interf", "ati_phase", "slc1_mag", "dpca_diff", "dpca_mag", "mag_mask", "ati_phase_masked"}
    if not expect <= targets:
        raise RuntimeError(f"reference inline GMTI block moved: found {sorted(t for t in targets if t)}")
    ns = {"np": np, "slc1": slc1, "slc2": slc2}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return {k: ns[k] for k in expect}


def vehicle_target_generators():
    """The five generators of vehicle_targets.py (importable as-is: numpy only)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_ref_vehicle_targets", os.path.join(REFERENCE_DIR, "vehicle_targets.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return {n: getattr(mod, n) for n in
            ("generate_car", "generate_tank", "generate_fighter_jet", "generate_f35", "generate_destroyer")}


def rda_functions():
    """``sar_focus_rda`` of the three simulators that define it (sar_satellite_sim.py:356-448 -> 7-tuple,
    sar_vehicle_sim.py:182-274 -> 8-tuple, sar_satellite_moving_sim.py:208-285 -> 3-tuple).  They read no module
    globals; scipy supplies interp1d / convolve / hamming as in the scripts' imports."""
    from scipy.interpolate import interp1d
    from scipy.signal import convolve
    from scipy.signal.windows import hamming
    out = {}
    for key, fname in (("satellite", "sar_satellite_sim.py"), ("vehicle", "sar_vehicle_sim.py"),
                       ("moving", "sar_satellite_moving_sim.py")):
        ns = {"np": np, "interp1d": interp1d, "convolve": convolve, "hamming": hamming}
        out[key] = _quiet(_extract(fname, ("sar_focus_rda",), ns)["sar_focus_rda"])
    return out


def _constants(filename: str, names: tuple[str, ...]) -> dict:
    """Literal module-level assignments (``P_TX = 1000.0``) of a reference script, by name."""
    path = os.path.join(REFERENCE_DIR, filename)
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) \
                and node.targets[0].id in names:
            out[node.targets[0].id] = ast.literal_eval(node.value)
    missing = set(names) - set(out)
    if missing:
        raise RuntimeError(f"{filename}: constants not found: {sorted(missing)}")
    return out


def noise_functions(fname: str = "sar_satellite_sim.py"):
    """``calculate_snr_db`` and ``add_ocean_noise`` (sar_satellite_sim.py:319-344; identical copies in
    sar_vehicle_sim.py:140-165 and sar_satellite_moving_sim.py:174-206 with their own radar constants)."""
    ns = {"np": np}
    ns.update(_constants(fname, ("P_TX", "ANT_LENGTH", "ANT_WIDTH", "T_SYS", "NF_DB", "LOSS_DB", "K_BOLTZ", "SCR_DB",
                                 "K_NU")))
    f = _extract(fname, ("calculate_snr_db", "add_ocean_noise"), ns)
    return f["calculate_snr_db"], _quiet(f["add_ocean_noise"])


def viewer_sardata_class():
    """The ``SARData`` class nested in ``main()`` of sar_ati_dcpa_viewer_csa.py (:35-55)."""
    path = os.path.join(REFERENCE_DIR, "sar_ati_dcpa_viewer_csa.py")
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    nodes = [n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "SARData"]
    if not nodes:
        raise RuntimeError("SARData not found")
    ns = {"np": np}
    exec(compile(ast.Module(body=[nodes[0]], type_ignores=[]), path, "exec"), ns)
    return ns["SARData"]


def batch_functions(g: dict):
    """``run_physics_spotlight`` and ``tdbp_gpu`` of sar_batch_sim.py (:83-169, :171-238) on the CPU, bound to the module
    constants in ``g`` (C, R0, FC, T_P, K_RATE, FS, Lambda)."""
    import torch
    ns = {"np": np, "torch": torch, "device": torch.device("cpu")}
    ns.update({k: g[k] for k in ("C", "R0", "FC", "T_P", "K_RATE", "FS", "Lambda")})
    f = _extract("sar_batch_sim.py", ("run_physics_spotlight", "tdbp_gpu"), ns)
    return _quiet(f["run_physics_spotlight"]), _quiet(f["tdbp_gpu"])


def batch_snr_function():
    """``calculate_raw_snr_db`` of sar_batch_sim.py:53-63 with its module constants."""
    ns = {"np": np}
    ns.update(_constants("sar_batch_sim.py", ("P_TX", "ANT_WIDTH", "T_SYS", "NF_DB", "LOSS_DB", "K_BOLTZ")))
    return _extract("sar_batch_sim.py", ("calculate_raw_snr_db",), ns)["calculate_raw_snr_db"]
