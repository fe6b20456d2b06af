"""Host side of the peer-mapped table (KGEkeras/_peer.py) that needs no GPU: the file-descriptor exchange between the ranks'
processes (UNIX sockets, SCM_RIGHTS) and the shard-size arithmetic."""
import multiprocessing as mp
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _fd_worker(rank, world, tag, sockdir, q):
    for p in (ROOT, os.path.join(ROOT, "kge-keras_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["KGEB_SOCKET_DIR"] = sockdir
    from KGEkeras._peer import exchange_fds
    r, w = os.pipe()
    os.write(w, f"hello from {rank}".encode())
    os.close(w)
    got = exchange_fds(rank, world, r, tag, timeout=30.0)
    out = {}
    for src, fd in got.items():
        out[src] = os.read(fd, 64).decode()        # the descriptor is live in THIS process: it reads the sender's pipe
        os.close(fd)
    q.put((rank, out))


@pytest.mark.parametrize("world", [2, 4])
def test_descriptor_exchange_between_processes(world, tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fd_worker, args=(r, world, f"t{world}", str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    # every pipe has one reader-side message: exactly one receiver gets the bytes, the others see EOF -- but every rank must
    # have received a descriptor from every other rank
    for r in range(world):
        assert sorted(res[r]) == [g for g in range(world) if g != r]
    for src in range(world):
        texts = [res[r][src] for r in range(world) if r != src]
        assert sorted(t for t in texts if t) == [f"hello from {src}"]
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".sock")]      # socket files are removed


def test_exchange_world1_is_empty():
    sys.path.insert(0, os.path.join(ROOT, "kge-keras_b200"))
    from KGEkeras._peer import exchange_fds
    assert exchange_fds(0, 1, 0, "x") == {}
