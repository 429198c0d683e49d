"""CPU-side checks of the boundary: the shared library loads without a GPU, exports every symbol
that include/nis_sar.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT

from nis_sar import _lib, params, scenes


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "nis_sar.h")).read()
    return sorted(set(re.findall(r"NIS_API\s+[a-z_0-9]+\s+(nis_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 16
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in nis_sar.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_struct_layouts():
    lib = _lib.load()
    assert lib.nis_version() == 2          # r2: caller-owned K3 workspace, nis_comm_*, nis_peer_native_atomics
    assert C.sizeof(_lib.EchoParams) == 56
    assert C.sizeof(_lib.CsaParams) == 64
    assert C.sizeof(_lib.GmtiResult) == 16
    # K3 workspace: 16-byte header + per 2048-pixel tile a 256-byte detection bitmap, a count and an arg-max candidate
    assert lib.nis_gmti_workspace_bytes(1) == 16 + 264 and lib.nis_gmti_workspace_bytes(4096 * 4096) == 16 + 264 * 8192


def test_graft_entry_build_passes():
    """The driver's "does it build" check: __graft_entry__.build() (make is a no-op when the library is current)."""
    import __graft_entry__ as g
    g.build()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    out = C.c_void_p()
    rc = lib.nis_ctx_create(0, C.byref(out))
    assert rc != 0 and not out.value
    assert "no CPU fallback" in _lib.last_error()
    from nis_sar import api
    with pytest.raises(Exception):
        api.sar_focus_csa(np.zeros((64, 64), np.complex64), 0.03, 20e-6, 2.5e13, 600e6, 6000.0, 7500.0, 5e5, 3.3e-3)


def test_collective_entry_points_without_a_device():
    """nis_comm_* bind NCCL at run time: without a GPU (or without NCCL) they fail with an error class and a message,
    they never crash the process; argument checks come first."""
    import torch  # noqa: F401  (loads libnccl.so.2 into the process when torch ships it)
    lib = _lib.load()
    assert lib.nis_comm_init(0, 0, None, None) == -1 and "bad argument" in _lib.last_error()
    assert lib.nis_echo_reduce(None, None, 10, -1, None) == -1
    assert lib.nis_slc_exchange(None, None, None, 10, None) == -1
    assert lib.nis_peer_native_atomics(0, 0) == 1
    assert lib.nis_comm_destroy(None) == 0


def test_install_patches_a_namespace():
    """nis_sar.install(): the wrappers keep the reference's names / positional signatures, inject live radar constants only
    into entry points that read them, and leave argument-only functions (sar_focus_csa, sar_focus_rda) unwrapped."""
    import inspect
    from nis_sar import api
    ns = {"C": 3e8, "R0": 5e5, "FC": 9.65e9, "BW": 5e8, "T_p": 2e-5, "FS": 6e8}
    assert api.install(ns) is ns
    for n in api._ENTRY_POINTS:
        assert callable(ns[n]) and ns[n].__name__ == n, n
    sig = inspect.signature(ns["run_bistatic_physics_gpu"])
    assert list(sig.parameters)[:6] == ["targets", "t_vec", "pos_tx_np", "vel_tx_np", "rx_offset_dist", "vel_target_np"]
    assert list(inspect.signature(ns["sar_focus_csa"]).parameters)[:9] == [
        "phist", "center_wavelength_m", "pulse_width_sec", "chirp_rate_hzpsec", "sample_rate_hz", "prf_hz",
        "platform_speed_mps", "range_ref_m", "t_start_fast"]
    assert ns["sar_focus_csa"].keywords == {"order": "F"}           # the reference's img.T view semantics (:396)
    ns2 = api.install({}, names=("sar_focus_rda", "calculate_snr_db"), rda_returns="vehicle", snr_preset="vehicle")
    assert ns2["sar_focus_rda"].keywords == {"returns": "vehicle"}
    snr, gain = ns2["calculate_snr_db"](2.8e4, 10.0, 0.03, 3.0e8, 16.384)
    assert np.isfinite(snr) and gain > 0
    with pytest.raises(Exception, match="not a replaceable"):
        api.install({}, names=("save_plot",))
    # without a GPU the patched functions raise instead of computing on the CPU
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            ns["run_custom_physics"]([{"position": [0, 0, 0], "rcs": 1.0}], np.zeros(2), np.zeros((2, 3)), 5e-4, 1e-6, 1e10, 3e8)


def test_hostio_placement_helpers():
    from nis_sar import hostio
    assert hostio._cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    nodes = hostio.numa_nodes()
    assert isinstance(nodes, dict)
    info = hostio.bind_rank_to_numa(0, 1, policy="none")
    assert info["node"] is None


def test_size_classes():
    lib = _lib.load()
    assert lib.nis_csa_size_class(4096, 4096) == 1
    assert lib.nis_csa_size_class(8192, 8192) == 1
    assert lib.nis_csa_size_class(64, 16384) == 1
    assert lib.nis_csa_size_class(32768, 2048) == 1       # the aperture sar_vehicle_sim.py focuses (:43)
    assert lib.nis_csa_size_class(48, 4096) == 2          # general-size path (mixed radix)
    assert lib.nis_csa_size_class(7199, 13200) == 2       # the reference's default scene
    assert lib.nis_csa_size_class(8209, 4096) == 0        # prime > 8192: no engine
    assert lib.nis_csa_size_class(1, 64) == 0


def test_presets_match_reference_constants():
    p = params.spaceborne_preset()
    # sar_ati_dcpa_sim_csa.py:18-46 evaluated by hand
    assert abs(p.V_sat - np.sqrt(3.986004418e14 / (6371000.0 + 350000.0))) < 1e-9
    assert abs(p.R0 - 509385.5) < 1.0 and abs(p.V_eff - 7497.9) < 0.5
    assert abs(p.d_rx - 2 * p.V_sat / 6000.0) < 1e-12
    assert p.num_samples == 13200 and p.t_start_fast == (2 * p.R0 / p.C) - 10e-6 - 1e-6
    v = params.airborne_vehicle_preset()
    assert v.num_samples == 2048 and v.k_rate == 300e6 / 1e-6


def test_scene_builders_are_seeded():
    a = scenes.ati_scene(seed=4, num_pulses=16, num_clutter=10)
    b = scenes.ati_scene(seed=4, num_pulses=16, num_clutter=10)
    c = scenes.ati_scene(seed=5, num_pulses=16, num_clutter=10)
    assert np.array_equal(a["clutter_pos"], b["clutter_pos"]) and not np.array_equal(a["clutter_pos"], c["clutter_pos"])
    assert a["pos_tx"].shape == (16, 3) and abs(np.linalg.norm(a["vel_tx"][0]) - a["prm"].V_sat) < 1e-6
    # broadside slant range is R0
    s = scenes.ati_scene(seed=0, num_pulses=3, num_clutter=0)
    assert abs(np.linalg.norm(s["pos_tx"][1]) - s["prm"].R0) < 1e-3
    pos, rcs = scenes.dense_vehicle_scene(1, 500)
    assert pos.shape == (500, 3) and rcs.shape == (500,)
    # HRWS-N: phase centres (k - (N-1)/2) d_rx; N = 2 is the reference's own pair (sar_ati_dcpa_sim_csa.py:184-196)
    h8 = scenes.hrws_scene(8, seed=4, num_pulses=16, num_clutter=10)
    d = h8["prm"].d_rx
    assert np.allclose(h8["rx_offsets"], [(k - 3.5) * d for k in range(8)], rtol=0, atol=1e-12)
    assert np.allclose(np.diff(h8["rx_offsets"]), d) and np.array_equal(h8["clutter_pos"], a["clutter_pos"])
    assert scenes.hrws_scene(2, seed=4, num_pulses=16, num_clutter=10)["rx_offsets"] == a["rx_offsets"]


def test_fft_engine_index_arithmetic_on_the_host():
    """fft.cuh's and mixed_ct.cuh's pass functions are __host__ __device__: csrc/hosttest runs the exact scatter / gather /
    twiddle index arithmetic of every plan the kernels instantiate (padded and strided shared-memory layouts included) on the
    CPU, threads one after another between the barriers, against a double-precision DFT."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    csrc = os.path.join(ROOT, "nis-sar-amtigmti-video_b200", "csrc")
    subprocess.run(["make", "-C", csrc, "hosttest"], check=True, capture_output=True, timeout=300)
    out = subprocess.run([os.path.join(csrc, "..", "lib", "test_fft_host")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("all ok"), out.stdout[-2000:]
    assert out.stdout.count("fwd") >= 19            # one line per plan
    # mixed_ct.cuh (13200 = 11.10.10.12, 7200 = 9.10.10.8): small DFTs, twiddle tables (full and powers-of-w^k), forward passes
    # 0..3 and the transposed inverse passes 3..0 in the order k_row_mixed_ct runs them
    out = subprocess.run([os.path.join(csrc, "..", "lib", "test_mixed_host")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("all ok"), out.stdout[-2000:]
    assert "MP13200" in out.stdout and "MP7200" in out.stdout


def test_videosar_timeline_matches_sar_batch_sim():
    """The sliding-CPI frame loop of sar_batch_sim.py (:244-252, :303-306) at its own constants (PRF 5000, 5 s, 10 fps,
    0.5 s CPI): 25000 pulses, frames every 500 pulses, 2500-pulse CPIs, and only 46 of the 50 requested frames fit."""
    from nis_sar import video
    prm = params.batch_spotlight_preset()
    t_all, step, cpi, nfr = video.batch_timeline(prm)
    assert (len(t_all), step, cpi, nfr) == (25000, 500, 2500, 50)
    assert t_all[0] == -2.5 and t_all[-1] == 2.5
    wins = video.cpi_windows(len(t_all), step, cpi, nfr)
    assert len(wins) == 46 and wins[0] == (0, 2500) and wins[-1] == (22500, 25000)
    assert video.cpi_windows(100, 10, 101, 5) == [] and video.cpi_windows(100, 10, 100, 5) == [(0, 100)]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys,
    the reference-arm additions, and the same metric / unit / config as the GPU arm."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, NIS_REF_BUDGET_S="20")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "csa_focused_mpixels_per_s" and d["unit"] == "Mpixel/s"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"].startswith("port") and "extrapolated" in d["cpu_baseline"]["kind"]
    assert d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"] and d["steps_run"] == 1
    assert "4096x4096" in d["cpu_baseline"]["sample"] and "64 of 8192 pulses" in d["cpu_baseline"]["sample"]


def test_host_transfer_thread_policy(monkeypatch):
    """nis_sar.hostio.host_threads(): NIS_HOST_THREADS wins; otherwise the cores of this process are divided among the
    ranks of the node and one is left to the thread that drives the DMA; the host-converted route needs >= 6 of them."""
    from nis_sar import hostio
    monkeypatch.setenv("NIS_HOST_THREADS", "12")
    assert hostio.host_threads() == 12 and hostio._host_route(1 << 24) == 12 and hostio._host_route(1000) == 0
    monkeypatch.setenv("NIS_HOST_THREADS", "0")
    assert hostio.host_threads() == 0 and hostio._host_route(1 << 24) == 0
    monkeypatch.delenv("NIS_HOST_THREADS")
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(32)), raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert hostio.host_threads() == 12
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert hostio.host_threads() == 2 and hostio._host_route(1 << 24) == 0      # 8 ranks on 32 cores: device route
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
    assert hostio.host_threads() == 12
    lib = _lib.load()
    assert lib.nis_d2h_widen(None, None, None, 4, 4, None) == -1 and "null argument" in _lib.last_error()
