"""TEST INFRASTRUCTURE: numpy emulation of the streaming kernels' index arithmetic
(mentpy_b200/csrc/stream.cuh), used to test the host-side pass scheduling of
mentpy_b200/streaming.py on CPU -- single process, and world_size > 1 over gloo where the peer
half that the CUDA kernel reads over NVLink is fetched with send/recv."""
import numpy as np


def _insert_fields(t, ranges):
    g = t.copy()
    for pos, wd in ranges:
        lo = g & ((np.uint64(1) << np.uint64(pos)) - np.uint64(1))
        g = ((g >> np.uint64(pos)) << np.uint64(pos + wd)) | lo
    return g


def _parity(x):
    x = x.copy()
    for s in (32, 16, 8, 4, 2, 1):
        x ^= x >> np.uint64(s)
    return (x & np.uint64(1)).astype(np.int64)


class NumpyStreamEngine:
    def __init__(self, dist=None):
        self.dist = dist  # torch.distributed module when world > 1

    def init(self, plan, L, rank, input_state, defer=False):  # the emulation always materialises the seed
        w = plan.window
        self.L, self.rank = L, rank
        n = 1 << L
        idx = (np.arange(n, dtype=np.uint64)) | np.uint64(rank << L)
        n_in = len(plan.input_slot)
        if input_state is None:
            psi = np.full(n, 2.0 ** (-w / 2), dtype=complex)
        else:
            src = np.zeros(n, dtype=np.int64)
            for q, sl in enumerate(plan.input_slot):
                src |= ((idx >> np.uint64(sl)) & np.uint64(1)).astype(np.int64) << (n_in - 1 - q)
            psi = np.asarray(input_state, dtype=complex)[src] * 2.0 ** (-(w - n_in) / 2)
        sg = np.zeros(n, dtype=np.int64)
        for a in range(w):
            bit = ((idx >> np.uint64(a)) & np.uint64(1)).astype(np.int64)
            sg ^= bit & _parity(idx & np.uint64(plan.init_cz_mask[a]))
        psi = psi * (1 - 2 * sg)
        half = n // 2
        self.H = [psi[:half].copy(), psi[half:].copy()]  # H0, H1 (top local bit)

    def _full(self):
        return np.concatenate(self.H)

    def _set_full(self, v):
        half = len(v) // 2
        self.H = [v[:half].copy(), v[half:].copy()]

    def local_pass(self, p, index_or, seeded=False):
        psi = self._full()
        K = len(p.slots)
        t = np.arange(p.n_groups, dtype=np.uint64)
        g = _insert_fields(t, p.ranges)
        ofs = [sum((1 << p.slots[j]) for j in range(K) if (l >> j) & 1) for l in range(1 << K)]
        a = [psi[(g + np.uint64(o)).astype(np.int64)] * p.scale for o in ofs]
        gfull = g | np.uint64(index_or)
        for j in range(K):
            e = p.cos_t[j] - 1j * p.sin_t[j]
            pg = _parity(gfull & np.uint64(p.nbr_masks[j]))
            for l in range(1 << K):
                if (l >> j) & 1:
                    continue
                lj = l | (1 << j)
                tt = a[l] + e * a[lj]
                a[l] = tt
                par = pg ^ (bin(l & p.local_masks[j]).count("1") & 1)
                a[lj] = tt * (1 - 2 * par)
        dead_local = sum(1 << j for j in range(K) if not (p.append_mask >> j) & 1)
        for l, o in enumerate(ofs):
            if l & dead_local == 0:
                psi[(g + np.uint64(o)).astype(np.int64)] = a[l]
        self._set_full(psi)

    def _peer_half(self, which_i_need, which_partner_needs, partner):
        """Fetch partner's half `which_i_need`, serving it my half `which_partner_needs`."""
        import torch

        mine = torch.from_numpy(np.ascontiguousarray(self.H[which_partner_needs]).view(np.float64).copy())
        theirs = torch.empty_like(mine)
        if self.rank < partner:
            self.dist.send(mine, partner)
            self.dist.recv(theirs, partner)
        else:
            self.dist.recv(theirs, partner)
            self.dist.send(mine, partner)
        return theirs.numpy().view(np.complex128)

    def exchange(self, p, role, partner, const_parity):
        e = p.cos_t - 1j * p.sin_t
        if role == 2:
            for h in (0, 1):
                peer = self._peer_half(h, h, partner)   # survivor reads; the dying side only serves
                self.H[h] = (self.H[h] + e * peer) * p.scale
            return
        peer = self._peer_half(role, 1 - role, partner)
        i = np.arange(len(peer), dtype=np.uint64)
        sgn = 1 - 2 * (const_parity ^ _parity(i & np.uint64(p.half_mask)))
        if role == 0:
            t = (self.H[0] + e * peer) * p.scale
            self.H = [t, t * sgn]
        else:
            t = (peer + e * self.H[1]) * p.scale
            self.H = [t, t * sgn]

    def serve_dying(self, p, partner):
        """role of the rank whose share dies in a tail exchange: only provides its halves."""
        for h in (0, 1):
            self._peer_half(h, h, partner)

    def rotate_roles(self, shard_bit):
        pass

    def gather(self, output_slots, L, rank, alive):
        k = len(output_slots)
        out = np.zeros(1 << k, dtype=complex)
        if alive:
            psi = self._full()
            for o in range(1 << k):
                idx = 0
                for q, sl in enumerate(output_slots):
                    idx |= ((o >> (k - 1 - q)) & 1) << sl
                if idx >> L == rank:
                    out[o] = psi[idx & ((1 << L) - 1)]
        return out

    def allreduce(self, vec):
        if self.dist is not None:
            import torch

            t = torch.from_numpy(vec.view(np.float64).copy())
            self.dist.all_reduce(t)
            return t.numpy().view(np.complex128)
        return vec

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
