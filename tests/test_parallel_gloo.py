"""CPU tests (gloo, world_size 2) of the host-side logic of the multi-GPU path: slab ranges, the layout of the
distributed-FFT transpose (send [peer][z_loc][ky_loc][kx][9], receive [z][ky_loc][kx][9]) and the all-to-all itself.
The per-rank transforms are done with numpy here -- the point is the decomposition and exchange indexing that
solver_cpp_b200.parallel and the library's pass address maps (csrc/fft_host.h: DmkLineMap) agree on."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from solver_cpp_b200 import parallel
    nx, ny, nz = 6, 4 * world, 2 * world
    nxr = nx // 2 + 1
    rng = np.random.default_rng(0)
    field = rng.standard_normal((nz, ny, nx, 9))                      # same on every rank
    z0, z1 = parallel.slab_range(nz, world, rank)
    send_shape, recv_shape = parallel.exchange_layout(nxr, ny, nz, world)
    # local x and y transforms of my z-slab
    loc = np.fft.fft(np.fft.rfft(field[z0:z1], axis=2), axis=1)       # [z_loc][ky][kx][9]
    ky_loc = ny // world
    send = np.stack([loc[:, p * ky_loc:(p + 1) * ky_loc] for p in range(world)])     # [peer][z_loc][ky_loc][kx][9]
    assert send.shape == send_shape
    st = torch.from_numpy(np.ascontiguousarray(np.stack([send.real, send.imag], -1)))
    rt = parallel.alltoall_transpose_host(st)
    recv = rt.numpy()[..., 0] + 1j * rt.numpy()[..., 1]
    recv = recv.reshape(recv_shape)                                    # [z][ky_loc][kx][9]
    hat = np.fft.fft(recv, axis=0)
    ref = np.fft.fftn(np.fft.rfft(field, axis=2), axes=(0, 1))[:, rank * ky_loc:(rank + 1) * ky_loc]
    err = np.abs(hat - ref).max() / np.abs(ref).max()
    # all-reduce of the partial sums every rank contributes (P_av, div-RMS)
    part = torch.tensor([field[z0:z1].sum()])
    dist.all_reduce(part)
    q.put((rank, float(err), float(part.item()), float(field.sum())))
    dist.destroy_process_group()


def test_slab_range_and_layout():
    sys.path.insert(0, ROOT)
    from solver_cpp_b200 import parallel, lib
    assert [parallel.slab_range(8, 4, r) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    with pytest.raises(lib.DmkError):
        parallel.slab_range(7, 2, 0)
    assert parallel.exchange_layout(5, 8, 4, 2) == ((2, 2, 4, 5, 9), (4, 4, 5, 9))


def test_fused_exchange_addressing_equals_the_all_to_all():
    """the store addresses of the fused exchange over peer memory (DmkPeerStore, csrc/spectral.cu) reproduce, rank by rank,
    exactly what the all-to-all transposes deliver: forward into the z-pass layout, inverse into the exchange layout"""
    sys.path.insert(0, ROOT)
    from solver_cpp_b200 import parallel
    for world, (nx, ny, nz) in ((2, (6, 8, 4)), (4, (4, 8, 12)), (8, (2, 16, 8))):
        nxr = nx // 2 + 1
        ny_loc, nz_loc = ny // world, nz // world
        rng = np.random.default_rng(world)
        full = rng.standard_normal((nz, ny, nxr))                       # one value per record, global (z, ky, kx)
        # forward: rank me holds [z_loc][ky][kx]; after the exchange rank p must hold [kz][ky_loc][kx] = full[:, p's ky range]
        bufB = [np.full(nz * ny_loc * nxr, np.nan) for _ in range(world)]
        for me in range(world):
            tgt = parallel.peer_store_forward(nxr, ny, nz, world, me)
            for z_loc in range(nz_loc):
                for ky in range(ny):
                    for kx in range(nxr):
                        p, rec = tgt(z_loc, ky, kx)
                        bufB[p][rec] = full[me * nz_loc + z_loc, ky, kx]
        for p in range(world):
            assert np.array_equal(bufB[p].reshape(nz, ny_loc, nxr), full[:, p * ny_loc:(p + 1) * ny_loc])
        # inverse: rank me holds [kz][ky_loc][kx]; rank p must receive [src][z_loc][ky_loc][kx] (what the inverse y pass reads)
        bufA = [np.full(world * nz_loc * ny_loc * nxr, np.nan) for _ in range(world)]
        for me in range(world):
            tgt = parallel.peer_store_inverse(nxr, ny, nz, world, me)
            for ky_loc in range(ny_loc):
                for kz in range(nz):
                    for kx in range(nxr):
                        p, rec = tgt(ky_loc, kz, kx)
                        bufA[p][rec] = full[kz, me * ny_loc + ky_loc, kx]
        for p in range(world):
            got = bufA[p].reshape(world, nz_loc, ny_loc, nxr)
            for src in range(world):
                assert np.array_equal(got[src], full[p * nz_loc:(p + 1) * nz_loc, src * ny_loc:(src + 1) * ny_loc])


def test_update_coords_halo_slots():
    """the two-slots-per-peer all-to-all of update_coords delivers, for every world size, the top layer of the rank below and
    the bottom layer of the rank above (for 2 ranks both slots of the one peer are used)"""
    sys.path.insert(0, ROOT)
    from solver_cpp_b200 import parallel
    for world in (2, 3, 4, 8):
        # all-to-all: what rank r puts in slot (p, h) arrives at rank p in slot (r, h)
        arrived = [dict() for _ in range(world)]
        for r in range(world):
            send, _ = parallel.halo_slots(world, r)
            for (p, h), which in send.items():
                assert (r, h) not in arrived[p]
                arrived[p][(r, h)] = (r, which)
        for r in range(world):
            _, recv = parallel.halo_slots(world, r)
            assert arrived[r][recv['above']] == ((r + 1) % world, 'bottom')
            assert arrived[r][recv['below']] == ((r - 1) % world, 'top')


def test_transpose_exchange_world_size_2():
    world, port = 2, 29533
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, err, tot, ref in res:
        assert err < 1e-13, (rank, err)
        assert abs(tot - ref) < 1e-9 * abs(ref)
