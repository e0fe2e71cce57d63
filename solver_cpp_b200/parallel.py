"""z-slab decomposition and the collective callbacks of the multi-GPU path (one process per GPU).

Decomposition as the reference's DiscretizationGrid (/root/reference/src/grid/discretization_grid.cpp:40-56, FFTW-MPI
default blocks): real space is split along z, Fourier space along y (FFTW_MPI_TRANSPOSED_OUT).  The library asks for
three collectives through the dmk_comm callback table (include/dmk_grid.h): the all-to-all of the distributed-FFT
transpose and sum / max all-reduces of a few doubles.  Here they are bound to torch.distributed (NCCL over
NVLink/NVSwitch on GPUs; gloo on CPU tensors for the host-logic tests)."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import lib as L


def slab_range(nz, world_size, rank):
    """contiguous z-range [z0, z1) of `rank`; requires nz % world_size == 0 like the library"""
    if nz % world_size:
        raise L.DmkError('nz = %d is not divisible by the number of ranks (%d)' % (nz, world_size))
    n = nz // world_size
    return rank * n, (rank + 1) * n


def exchange_layout(nxr, ny, nz, world_size):
    """shapes of the transpose buffers: send [peer][z_loc][ky_loc][kx][9] complex, received [z][ky_loc][kx][9]"""
    return (world_size, nz // world_size, ny // world_size, nxr, 9), (nz, ny // world_size, nxr, 9)


class _DevPtr:
    """zero-copy view of raw device memory for torch.as_tensor (CUDA array interface v2)"""

    def __init__(self, ptr, n_doubles):
        self.__cuda_array_interface__ = {'shape': (n_doubles,), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}


def _view(ptr, n_doubles, device):
    return torch.as_tensor(_DevPtr(ptr, n_doubles), device=device)


class TorchComm:
    """dmk_comm callbacks on torch.distributed; keeps the ctypes thunks alive"""

    def __init__(self, world_size, rank, group=None):
        self.world_size, self.rank, self.group = world_size, rank, group
        self.device = torch.device('cuda', torch.cuda.current_device())

        def alltoall(user, d_send, d_recv, bytes_per_peer):
            try:
                n = bytes_per_peer // 8 * self.world_size
                send, recv = _view(d_send, n, self.device), _view(d_recv, n, self.device)
                dist.all_to_all_single(recv, send, group=self.group)
                return 0
            except Exception as e:                                   # surfaced as rc != 0 in the library
                print('alltoall callback failed:', e)
                return 1

        def allreduce(op):
            def f(user, d_buf, n):
                try:
                    dist.all_reduce(_view(d_buf, n, self.device), op=op, group=self.group)
                    return 0
                except Exception as e:
                    print('allreduce callback failed:', e)
                    return 1
            return f
        self._a2a = L.ALLTOALL_FN(alltoall)
        self._sum = L.ALLREDUCE_FN(allreduce(dist.ReduceOp.SUM))
        self._max = L.ALLREDUCE_FN(allreduce(dist.ReduceOp.MAX))
        self.struct = L.Comm(None, self._a2a, self._sum, self._max)


class NcclComm:
    """the library's own NCCL binding (csrc/nccl_comm.cpp): the collectives are issued by the library on its stream, no Python on
    the per-evaluation path.  torch.distributed only carries the 128-byte ncclUniqueId from rank 0 to the others, once."""

    def __init__(self, world_size, rank, group=None):
        self.world_size, self.rank, self.group = world_size, rank, group
        self.native = False
        self._fallback = None

    def attach(self, g):
        lib = g.lib
        uid = (C.c_ubyte * 128)()
        ok = 1
        try:
            if self.rank == 0:
                L.check(lib.dmk_nccl_unique_id(uid))
        except L.DmkError as e:
            print('solver_cpp_b200: native NCCL unavailable (%s)' % e)
            ok = 0
        t = torch.tensor(list(bytes(uid)) + [ok], dtype=torch.uint8, device=g.device)
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        dist.broadcast(t, src=src, group=self.group)
        vals = t.cpu().tolist()
        if vals[128]:
            try:
                L.check(lib.dmk_nccl_init(g.ctx, (C.c_ubyte * 128)(*vals[:128])))
                self.native = True
                return
            except L.DmkError as e:
                print('solver_cpp_b200: dmk_nccl_init failed on rank %d (%s)' % (self.rank, e))
                raise
        self._fallback = TorchComm(self.world_size, self.rank, self.group)       # every rank alike: rank 0's flag decides
        L.check(lib.dmk_set_comm(g.ctx, C.byref(self._fallback.struct)))


def make_comm(world_size, rank, group=None):
    """the collectives of a z-slab job: the library's native NCCL binding on GPUs (DMK_COMM=torch: the torch.distributed callback
    table instead; always that one for CPU / gloo groups, which only the host-logic tests use)"""
    import os
    native = os.environ.get('DMK_COMM', 'nccl').lower() != 'torch' and torch.cuda.is_available() and dist.get_backend(group) == 'nccl'
    return NcclComm(world_size, rank, group) if native else TorchComm(world_size, rank, group)


def fused_exchange_wanted():
    """DMK_EXCHANGE=nccl keeps the transposes on the NCCL all-to-all (comparison runs); default: fused over peer memory"""
    import os
    return os.environ.get('DMK_EXCHANGE', 'p2p').lower() != 'nccl'


def alltoall_transpose_host(send, group=None):
    """the transpose exchange on a CPU tensor (gloo): send [peer][...] -> recv [peer][...]; host-logic twin of the
    device path used by the world_size-2 tests"""
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.contiguous().view(-1), group=group) if dist.get_backend(group) != 'gloo' else \
        _gloo_alltoall(recv, send, group)
    return recv


def _gloo_alltoall(recv, send, group):
    # gloo has no all_to_all_single for CPU tensors in every build: emulate with P scatters (test sizes only)
    P = dist.get_world_size(group)
    me = dist.get_rank(group)
    for src in range(P):
        out = recv[src]
        dist.scatter(out, [send[p].contiguous() for p in range(P)] if me == src else None, src=src, group=group)
    return recv


def peer_store_forward(nxr, ny, nz, world_size, me):
    """host twin of the library's DmkPeerStore for the forward exchange (csrc/spectral.cu: peers_forward): where rank `me`
    stores y-line position ky of its local plane z_loc, record kx -> (peer, record index in the peer's z-pass buffer
    [kz][ky_loc][kx])"""
    ny_loc, nz_loc = ny // world_size, nz // world_size
    hi, o, e = nz_loc * ny_loc * nxr, ny_loc * nxr, nxr

    def target(z_loc, ky, kx):
        return ky // ny_loc, me * hi + z_loc * o + (ky % ny_loc) * e + kx
    return target


def peer_store_inverse(nxr, ny, nz, world_size, me):
    """... and for the inverse exchange (peers_inverse): z-line position kz of local ky_loc -> (peer, record index in the
    peer's y-pass buffer in the exchange layout [src][z_loc][ky_loc][kx])"""
    ny_loc, nz_loc = ny // world_size, nz // world_size
    hi, o, e = nz_loc * ny_loc * nxr, nxr, ny_loc * nxr

    def target(ky_loc, kz, kx):
        return kz // nz_loc, me * hi + ky_loc * o + (kz % nz_loc) * e + kx
    return target


def halo_slots(world_size, rank):
    """host twin of the z-halo exchange of update_coords (csrc/spectral.cu: dmk_update_coords): the all-to-all carries two
    layer slots per peer.  Returns (send, recv): send[(peer, half)] = 'bottom' | 'top' (which of my layers goes there),
    recv['above'] / recv['below'] = (peer, half) where the neighbouring layers arrive."""
    rb, rt = (rank - 1 + world_size) % world_size, (rank + 1) % world_size
    send = {(rb, 0): 'bottom', (rt, 1): 'top'}
    recv = {'above': (rt, 0), 'below': (rb, 1)}
    return send, recv
