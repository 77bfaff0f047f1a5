"""Sorted-interval set algebra on the device (``pb_chain_union`` / ``pb_chain_binary``): the position-set
arithmetic of ``cs generate`` (plastid/bin/cs.py:242-496) for all genes and transcripts of an
annotation at once.  A :class:`ChainSet` is a device-resident chain table: chain c owns the sorted,
disjoint, non-touching blocks ``[off[c], off[c+1])`` in global-bin coordinates."""
import numpy as np

from . import _lib


class ChainSet(object):
    def __init__(self, bstart, bend, off):
        self.bstart, self.bend, self.off = bstart, bend, off      # torch int64 tensors on one device

    @property
    def n_chains(self):
        return int(self.off.numel()) - 1

    @property
    def device(self):
        return self.off.device

    @classmethod
    def from_numpy(cls, bstart, bend, off, device):
        import torch

        def up(a):
            a = np.ascontiguousarray(a, dtype=np.int64)
            return torch.from_numpy(a if len(a) else np.zeros(1, dtype=np.int64)).to(device)
        return cls(up(bstart), up(bend), up(off))

    @staticmethod
    def cat(sets):
        """One table holding the chains of ``sets`` one after the other."""
        import torch
        nb = [int(s.off[-1].item()) for s in sets]
        bstart = torch.cat([s.bstart[:n] for s, n in zip(sets, nb)] + [torch.zeros(1, dtype=torch.int64, device=sets[0].device)])
        bend = torch.cat([s.bend[:n] for s, n in zip(sets, nb)] + [torch.zeros(1, dtype=torch.int64, device=sets[0].device)])
        offs, base = [sets[0].off[:1]], 0
        for s, n in zip(sets, nb):
            offs.append(s.off[1:] + base)
            base += n
        return ChainSet(bstart, bend, torch.cat(offs))

    def numpy(self):
        off = self.off.cpu().numpy()
        n = int(off[-1])
        return self.bstart[:n].cpu().numpy(), self.bend[:n].cpu().numpy(), off


def _two_pass(n_out, launch, device):
    import torch
    n_blk = torch.zeros(max(n_out, 1), dtype=torch.int32, device=device)
    launch(n_blk, None, None, None)
    off = torch.zeros(n_out + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(n_blk[:n_out].to(torch.int64), 0)
    total = int(off[-1].item())
    bstart = torch.zeros(max(total, 1), dtype=torch.int64, device=device)
    bend = torch.zeros(max(total, 1), dtype=torch.int64, device=device)
    if total:
        launch(n_blk, off, bstart, bend)
    return ChainSet(bstart, bend, off)


def chain_union(chains, grp_off, members):
    """Output chain g = union of ``chains[members[grp_off[g]:grp_off[g+1]]]``."""
    import torch
    _lib.require_cuda()
    dev = chains.device
    grp_off = np.ascontiguousarray(grp_off, dtype=np.int64)
    members = np.ascontiguousarray(members, dtype=np.int64)
    if len(members) and (members.min() < 0 or members.max() >= chains.n_chains):
        raise IndexError("chain_union: member index outside the chain table")
    n = len(grp_off) - 1
    d_off = torch.from_numpy(grp_off).to(dev)
    d_mem = torch.from_numpy(members if len(members) else np.zeros(1, dtype=np.int64)).to(dev)

    def launch(n_blk, off, bstart, bend):
        _lib.check(_lib.lib().pb_chain_union(_lib.ptr(chains.bstart), _lib.ptr(chains.bend), _lib.ptr(chains.off),
                                             _lib.ptr(d_off), _lib.ptr(d_mem), n, _lib.ptr(n_blk), _lib.ptr(off),
                                             _lib.ptr(bstart), _lib.ptr(bend), _lib.stream_ptr()))
    return _two_pass(n, launch, dev)


def chain_binary(op, a, a_idx, b, b_idx):
    """Output chain i = ``a[a_idx[i]]`` AND / SUB ``b[b_idx[i]]`` (``op`` "and" | "sub"; ``b_idx[i] < 0``
    = empty right-hand side)."""
    import torch
    _lib.require_cuda()
    dev = a.device
    a_idx = np.ascontiguousarray(a_idx, dtype=np.int64)
    b_idx = np.ascontiguousarray(b_idx, dtype=np.int64)
    if len(a_idx) != len(b_idx):
        raise ValueError("chain_binary: index arrays differ in length")
    if len(a_idx) and (a_idx.min() < 0 or a_idx.max() >= a.n_chains or b_idx.max() >= b.n_chains):
        raise IndexError("chain_binary: chain index outside its table")
    n = len(a_idx)
    code = {"and": _lib.PB_CHAIN_AND, "sub": _lib.PB_CHAIN_SUB}[op]
    d_a = torch.from_numpy(a_idx if n else np.zeros(1, dtype=np.int64)).to(dev)
    d_b = torch.from_numpy(b_idx if n else np.zeros(1, dtype=np.int64)).to(dev)

    def launch(n_blk, off, bstart, bend):
        _lib.check(_lib.lib().pb_chain_binary(code, _lib.ptr(a.bstart), _lib.ptr(a.bend), _lib.ptr(a.off), _lib.ptr(d_a),
                                              _lib.ptr(b.bstart), _lib.ptr(b.bend), _lib.ptr(b.off), _lib.ptr(d_b),
                                              n, _lib.ptr(n_blk), _lib.ptr(off), _lib.ptr(bstart), _lib.ptr(bend),
                                              _lib.stream_ptr()))
    return _two_pass(n, launch, dev)
