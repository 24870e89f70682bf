"""N > 1 host logic on CPU: frame -> rank map and the recon all-gather, world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpulib import pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nframes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = __import__("importlib").import_module("x265-mod-by-patman_b200.sharding")
    mine = sh.frames_for_rank(nframes, rank, world)
    plane = 1000
    gathered = []
    ok = True
    for step in range((nframes + world - 1) // world):
        f = step * world + rank
        recon = torch.full((plane,), f if f < nframes else -1, dtype=torch.int16)
        g = sh.exchange_recon(recon, world)
        gathered.append(g)
        for r in range(world):
            want = step * world + r
            ok &= bool((g[r] == (want if want < nframes else -1)).all())
    # every frame this rank owns can find its two previous reconstructions
    for f in mine:
        refs = sh.reference_planes(f, 2, world, gathered)
        ok &= len(refs) == min(2, f)
        for k, p in enumerate(refs, 1):
            ok &= int(p[0]) == f - k
    q.put((rank, mine, ok))
    dist.destroy_process_group()


def test_frame_sharding_and_recon_exchange_world2():
    world, nframes = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(f for _, mine, _ in res for f in mine)
    assert owned == list(range(nframes))          # every frame has exactly one owner
    assert all(ok for _, _, ok in res)


def test_owner_map():
    sh = __import__("importlib").import_module("x265-mod-by-patman_b200.sharding")
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            fr = sh.frames_for_rank(19, r, world)
            assert all(sh.owner_of(f, world) == r for f in fr)
            seen += fr
        assert sorted(seen) == list(range(19))
