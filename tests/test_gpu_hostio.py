"""Host <-> device transfers of the complex128 array contract (nis_sar.hostio, csrc/hostcopy.cpp): the route that moves
complex64 over PCIe and converts on the host cores must deliver the same BYTES as the device-side conversion (and as
numpy's astype), for every size class -- shorter than one chunk, ragged tails, many chunks, misaligned destinations."""
import numpy as np
import pytest
import torch

from nis_sar import hostio

pytestmark = pytest.mark.gpu

CHUNK = 1 << 17     # complex elements per transfer chunk (1 MiB of complex64)


def _rand_c64(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((n, 2), generator=g, device="cuda") * torch.tensor([1.0, 1e-20], device="cuda")
    return torch.view_as_complex(x.contiguous())


@pytest.mark.parametrize("threads", [1, 3, 8])
@pytest.mark.parametrize("n", [1, 7, 4099, CHUNK - 1, CHUNK + 5, 9 * CHUNK + 12345, 45 * CHUNK + 3])   # the last wraps every ring
def test_d2h_widen_is_bitwise_the_device_route(n, threads, monkeypatch):
    x = _rand_c64(n, n % 1000 + threads)
    monkeypatch.setenv("NIS_HOST_THREADS", "0")
    ref = hostio.to_host_c128(x).copy()
    assert hostio.last_transfer["d2h"]["threads"] == 0
    monkeypatch.setenv("NIS_HOST_THREADS", str(threads))
    monkeypatch.setattr(hostio, "_MIN_THREADS_FOR_HOST_ROUTE", 1)
    monkeypatch.setattr(hostio, "_MIN_ELEMENTS_FOR_HOST_ROUTE", 0)
    got = hostio.to_host_c128(x)
    assert hostio.last_transfer["d2h"]["threads"] == threads and hostio.last_transfer["d2h"]["pcie_bytes"] == 8 * n
    assert got.dtype == np.complex128 and got.shape == (n,)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))
    # pageable destination at a 16-byte (not 32-byte) offset, 2-D shape
    if n % 7 == 0 or n > CHUNK:
        big = np.full(n + 1, np.nan + 0j, dtype=np.complex128)
        out = big[1:]
        r = hostio.to_host_c128(x, out=out)
        assert r is out and np.array_equal(out.view(np.uint64), ref.view(np.uint64)) and np.isnan(big[0].real)


@pytest.mark.parametrize("threads", [1, 5])
@pytest.mark.parametrize("n", [3, 4099, CHUNK + 5, 6 * CHUNK + 777, 31 * CHUNK + 1])
def test_h2d_narrow_is_bitwise_numpy_astype(n, threads, monkeypatch):
    rng = np.random.default_rng(n + threads)
    h = (rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n) + 1j * rng.standard_normal(n)).astype(np.complex128)
    monkeypatch.setenv("NIS_HOST_THREADS", str(threads))
    monkeypatch.setattr(hostio, "_MIN_THREADS_FOR_HOST_ROUTE", 1)
    monkeypatch.setattr(hostio, "_MIN_ELEMENTS_FOR_HOST_ROUTE", 0)
    x = hostio.to_device_c64(h, "cuda:0")
    assert hostio.last_transfer["h2d"]["threads"] == threads
    want = h.astype(np.complex64)
    assert np.array_equal(x.cpu().numpy().view(np.uint32), want.view(np.uint32))
    monkeypatch.setenv("NIS_HOST_THREADS", "0")
    y = hostio.to_device_c64(h, "cuda:0")           # device-side narrowing: same rounding
    assert torch.equal(x.view(torch.float32), y.view(torch.float32))


def test_focus_returns_identical_arrays_on_both_routes(monkeypatch):
    """sar_focus_csa end to end (numpy complex128 in, complex128 out, C and F order): the transfer route is invisible."""
    from nis_sar import api, params
    prm = params.spaceborne_preset()
    rng = np.random.default_rng(5)
    raw = (rng.standard_normal((1024, 2048)) + 1j * rng.standard_normal((1024, 2048))).astype(np.complex128)
    args = (prm.Lambda, prm.T_p, prm.k_rate, prm.FS, prm.PRF, prm.V_eff, prm.R0, prm.t_start_fast)
    outs = {}
    for tag, thr in (("device", "0"), ("host", "6")):
        monkeypatch.setenv("NIS_HOST_THREADS", thr)
        monkeypatch.setattr(hostio, "_MIN_ELEMENTS_FOR_HOST_ROUTE", 0)
        outs[tag] = [api.sar_focus_csa(raw, *args, order=o)[0] for o in ("C", "F")]
    for a, b in zip(outs["device"], outs["host"]):
        assert a.flags == b.flags or (a.flags.f_contiguous == b.flags.f_contiguous)
        assert np.array_equal(a, b)
