"""N>1 host logic on CPU: world_size-2 gloo runs of the partition / collective wrappers in
nis_sar.dist, with the numpy oracle standing in for the CUDA operators (the wrappers take the
per-rank compute step as a callable)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sar_oracle as orc
from nis_sar import dist as nd, params, scenes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    prm = params.spaceborne_preset(fs=30e6, bw=25e6)      # S = 660
    sc = scenes.ati_scene(seed=8, num_pulses=7, num_clutter=25, prm=prm)
    pos = np.concatenate([sc["ship_pos"], sc["clutter_pos"]])
    rcs = np.concatenate([sc["ship_rcs"], sc["clutter_rcs"]])
    return prm, sc, pos, rcs


def _echo(prm, sc, pos, rcs, rows=slice(None)):
    raw, _ = orc.echo_bistatic(pos, rcs, sc["t_vec"][rows], sc["pos_tx"][rows], sc["vel_tx"][rows], sc["rx_offsets"][0],
                               np.zeros(3), prm.as_globals())
    return torch.from_numpy(raw.astype(np.complex64))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prm, sc, pos, rcs = _scene()
        full = _echo(prm, sc, pos, rcs)
        P, S = full.shape

        # scatterer shards + all-reduce
        summed = nd.echo_scatterer_shards(lambda t0, t1: _echo(prm, sc, pos[t0:t1], rcs[t0:t1]), len(rcs))
        e1 = float(torch.linalg.vector_norm(summed - full) / torch.linalg.vector_norm(full))
        # reduce to rank 0 only
        red = nd.echo_scatterer_shards(lambda t0, t1: _echo(prm, sc, pos[t0:t1], rcs[t0:t1]), len(rcs), dst=0)
        e1b = float(torch.linalg.vector_norm(red - full) / torch.linalg.vector_norm(full)) if rank == 0 else 0.0

        # pulse blocks, then gather
        raw = torch.zeros((P, S), dtype=torch.complex64)

        def fill(p0, p1, out):
            out[p0:p1] = _echo(prm, sc, pos, rcs, slice(p0, p1))
        nd.echo_pulse_blocks(fill, raw, gather=False)
        p0, p1 = nd.block_range(P, rank, world)
        own_ok = bool(torch.equal(raw[p0:p1], full[p0:p1])) and not bool(raw[:p0].any()) and not bool(raw[p1:].any())
        nd.echo_pulse_blocks(fill, raw, gather=True)
        e2 = float(torch.linalg.vector_norm(raw - full) / torch.linalg.vector_norm(full))

        # frame ownership
        frames = list(nd.frame_indices(5))
        got = nd.focus_frames(lambda f: torch.full((2, 2), float(f)), 5)

        # channel pairing: rank k receives channel k+1
        mine = torch.full((3, 4), complex(rank + 1, -rank), dtype=torch.complex64)
        other = nd.exchange_with_next(mine)
        pair = nd.pair_products(mine, lambda a, b: {"diff": a - b})
        tmax = nd.max_over_ranks(10.0 + rank, "cpu")
        q.put((rank, e1, e1b, own_ok, e2, frames, sorted(got), None if other is None else complex(other[0, 0]),
               None if pair is None else complex(pair["diff"][0, 0]), tmax))
    finally:
        dist.destroy_process_group()


def test_block_range_partitions():
    for n in (0, 1, 7, 8, 100000):
        for w in (1, 2, 3, 8):
            spans = [nd.block_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_fallthrough():
    # without an initialised process group every wrapper degenerates to the one-rank case
    assert nd.world() == (0, 1)
    assert list(nd.frame_indices(3)) == [0, 1, 2]
    t = torch.ones(2, 2, dtype=torch.complex64)
    assert nd.exchange_with_next(t) is None
    out = nd.echo_scatterer_shards(lambda a, b: torch.full((1,), float(b - a)), 10)
    assert float(out) == 10.0


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e1b, own_ok, e2, frames, got, other, pair, tmax in res:
        assert e1 < 1e-6 and e1b < 1e-6          # fp32 partial sums in a different order
        assert own_ok and e2 == 0.0
        assert frames == list(range(rank, 5, 2)) and got == frames
        assert tmax == 11.0
        if rank == 0:
            assert other == complex(2, -1) and pair == complex(1, 0) - complex(2, -1)
        else:
            assert other is None and pair is None
