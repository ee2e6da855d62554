"""Halo exchange for brick-partitioned box meshes: the PetscSF stand-in behind
DMGlobalToLocal(INSERT_VALUES) / DMLocalToGlobal(ADD_VALUES) (/root/reference/src/matops.c:33,57).

DMPlexDistribute with overlap 0 (/root/reference/src/setupdm.c:58-64) partitions ELEMENTS; mesh
nodes on partition interfaces are duplicated in the local vectors and owned by exactly one rank.
Here the partition is a grid of bricks and an interface node is owned by the LOWEST rank that
holds it.  Two collective steps, both neighbour point-to-point (<= 26 neighbours, 7 for 2x2x2):

  owner_to_ghost(Xloc)      owners send interface values, ghosts overwrite   (G2L, INSERT)
  ghost_to_owner_add(Yloc)  ghosts send partial sums, owners accumulate       (L2G, ADD)

Interface dofs are packed / unpacked with libceed_b200.so kernels (b200_gather,
b200_scatter_set, b200_scatter_add) and moved with torch.distributed batch_isend_irecv
(NCCL over NVLink for CUDA tensors, gloo for the CPU tests).
"""
import numpy as np
import torch


def _ranges(mesh_n, grid, p):
    """Per axis: node index range [lo, hi] (inclusive, global numbering) of every brick."""
    out = []
    for d in range(3):
        q, rem = divmod(mesh_n[d], grid[d])
        r = []
        for i in range(grid[d]):
            start = i * q + min(i, rem)
            size = q + (1 if i < rem else 0)
            r.append((start * p, (start + size) * p))
        out.append(r)
    return out


class Halo:
    def __init__(self, gmesh, grid, rank, p, dist=None, device=None, ncomp=3):
        self.dist, self.rank, self.grid, self.p = dist, rank, grid, p
        px, py, pz = grid
        rng = _ranges(gmesh.n, grid, p)
        me = (rank % px, (rank // px) % py, rank // (px * py))
        lo = [rng[d][me[d]][0] for d in range(3)]
        hi = [rng[d][me[d]][1] for d in range(3)]
        N = [hi[d] - lo[d] + 1 for d in range(3)]  # local nodes per axis
        self.nnodes = N[0] * N[1] * N[2]
        owner = np.full((N[2], N[1], N[0]), rank, dtype=np.int64)
        # shared[r] = local node indices (in a canonical GLOBAL order) shared with neighbour r
        self.neighbours = []
        shared = {}
        lidx = np.arange(self.nnodes).reshape(N[2], N[1], N[0])
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    if dx == dy == dz == 0:
                        continue
                    nb = (me[0] + dx, me[1] + dy, me[2] + dz)
                    if not all(0 <= nb[d] < grid[d] for d in range(3)):
                        continue
                    r = nb[0] + px * (nb[1] + py * nb[2])
                    sl = []
                    for d, dd in enumerate((dx, dy, dz)):
                        # overlap of my node range with the neighbour's along axis d
                        a = max(lo[d], rng[d][nb[d]][0]); b = min(hi[d], rng[d][nb[d]][1])
                        assert a <= b
                        sl.append(slice(a - lo[d], b - lo[d] + 1))
                    ids = lidx[sl[2], sl[1], sl[0]].reshape(-1)  # z,y,x order == global lexicographic order
                    shared[r] = ids
                    sub = owner[sl[2], sl[1], sl[0]]
                    np.minimum(sub, r, out=sub)
                    self.neighbours.append(r)
        self.neighbours.sort()
        self.owned_node_mask = (owner == rank).reshape(-1)
        own = owner.reshape(-1)
        dev = device if device is not None else ("cuda" if torch.cuda.is_available() and dist is not None and dist.get_backend() == "nccl" else "cpu")
        self.device = torch.device(dev)

        def dofs(nodes):
            return (nodes[:, None] * ncomp + np.arange(ncomp)[None, :]).reshape(-1).astype(np.int32)

        # owner -> ghost: I send the nodes I own that r holds; I receive the nodes r owns
        # ghost -> owner: the reverse.  Both ranks enumerate a shared set in the same global order.
        self.send_own, self.recv_ghost = {}, {}
        for r in self.neighbours:
            ids = shared[r]
            mine = ids[own[ids] == rank]
            theirs = ids[own[ids] == r]
            if mine.size:
                self.send_own[r] = torch.from_numpy(dofs(mine)).to(self.device)
            if theirs.size:
                self.recv_ghost[r] = torch.from_numpy(dofs(theirs)).to(self.device)
        # one contiguous pack buffer per direction: a single gather kernel packs the data for ALL
        # neighbours, a single scatter kernel unpacks it; sends/receives are views into the buffers
        def layout(idx):
            ranks = [r for r in self.neighbours if r in idx]
            cat = torch.cat([idx[r] for r in ranks]) if ranks else torch.zeros(0, dtype=torch.int32, device=self.device)
            views, o = {}, 0
            for r in ranks:
                views[r] = (o, o + idx[r].numel())
                o += idx[r].numel()
            return cat.contiguous(), views

        self.own_cat, self.own_views = layout(self.send_own)
        self.ghost_cat, self.ghost_views = layout(self.recv_ghost)
        # "sum-and-share": every dof shared with r, both directions at once (same order on both sides)
        self.shared_all = {r: torch.from_numpy(dofs(shared[r])).to(self.device) for r in self.neighbours}
        self.share_cat, self.share_views = layout(self.shared_all)
        # number of ranks holding each local node (1 in the interior)
        cnt = np.ones(self.nnodes)
        for r in self.neighbours:
            cnt[shared[r]] += 1
        self.rank_multiplicity = cnt
        self._buf = {}

    def _buffer(self, key, n, like):
        k = (key, like.dtype, like.device)
        if k not in self._buf:
            self._buf[k] = torch.empty(n, dtype=like.dtype, device=like.device)
        return self._buf[k]

    @staticmethod
    def _gather(dst, src, idx):
        if idx.numel() == 0:
            return
        if src.is_cuda:
            from .ceed import b2, lib
            b2(lib.b200_gather(dst.data_ptr(), src.data_ptr(), idx.data_ptr(), idx.numel()))
        else:
            torch.index_select(src, 0, idx.long(), out=dst)

    @staticmethod
    def _scatter(dst, idx, src, add):
        if idx.numel() == 0:
            return
        if dst.is_cuda:
            from .ceed import b2, lib
            f = lib.b200_scatter_add if add else lib.b200_scatter_set
            b2(f(dst.data_ptr(), idx.data_ptr(), src.data_ptr(), idx.numel()))
        elif add:
            dst.index_add_(0, idx.long(), src)
        else:
            dst.index_copy_(0, idx.long(), src)

    def _exchange(self, vec, send_cat, send_views, recv_cat, recv_views, add, tag):
        dist = self.dist
        sbuf = self._buffer("s" + tag, send_cat.numel(), vec)
        rbuf = self._buffer("r" + tag, recv_cat.numel(), vec)
        self._gather(sbuf, vec, send_cat)
        ops = []
        for r in self.neighbours:  # same global order on every rank
            if r in recv_views:
                a, b = recv_views[r]
                ops.append(dist.P2POp(dist.irecv, rbuf[a:b], r))
            if r in send_views:
                a, b = send_views[r]
                ops.append(dist.P2POp(dist.isend, sbuf[a:b], r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self._scatter(vec, recv_cat, rbuf, add)

    # ---- sum-and-share over NVLink peer memory (one node): no communication library on the data path
    def enable_p2p(self, timeout_s=5.0):
        """Collective over all ranks: allocate this rank's receive window, exchange CUDA IPC handles and segment
        layouts through torch.distributed (set-up only), map the neighbours' windows.  Afterwards sum_and_share
        is three kernels of libceed_b200.so (csrc/b200_halo.cu): store the packed partial sums into the
        neighbours' windows, flag them, wait for the neighbours' flags and add the own window."""
        import ctypes as C
        from .ceed import b2, lib
        dist = self.dist
        assert self.device.type == "cuda", "peer-memory halo needs CUDA tensors"
        total = int(self.share_cat.numel())
        nflags = 64
        buf = C.c_void_p()
        nbytes = 2 * total * 8 + nflags * 8 + 8
        b2(lib.b200_malloc(C.byref(buf), nbytes))
        b2(lib.b200_memset(buf, 0, nbytes))
        b2(lib.b200_sync())
        handle = (C.c_ubyte * 64)()
        b2(lib.b200_ipc_get_handle(buf, handle))
        mine = {"rank": self.rank, "handle": bytes(handle), "total": total, "neighbours": list(self.neighbours),
                "views": {r: self.share_views[r] for r in self.neighbours}}
        infos = [None] * dist.get_world_size()
        dist.all_gather_object(infos, mine)
        nn = len(self.neighbours)
        seg_start = (C.c_int * (nn + 1))()
        remote = [(C.c_void_p * max(nn, 1))(), (C.c_void_p * max(nn, 1))()]   # per parity
        rflag = (C.c_void_p * max(nn, 1))()
        self._p2p_open = []
        for s, r in enumerate(self.neighbours):
            a, b = self.share_views[r]
            seg_start[s], seg_start[s + 1] = a, b
            info = infos[r]
            ra, rb = info["views"][self.rank]
            assert rb - ra == b - a, "shared sets of a rank pair must have the same size on both sides"
            base = C.c_void_p()
            b2(lib.b200_ipc_open((C.c_ubyte * 64).from_buffer_copy(info["handle"]), C.byref(base)))
            self._p2p_open.append(base)
            for par in (0, 1):
                remote[par][s] = base.value + (par * info["total"] + ra) * 8
            rflag[s] = base.value + 2 * info["total"] * 8 + info["neighbours"].index(self.rank) * 8
        self._p2p = dict(buf=buf, total=total, nn=nn, seg_start=seg_start, remote=remote, rflag=rflag, gen=0,
                         flags=buf.value + 2 * total * 8, err=buf.value + 2 * total * 8 + nflags * 8,
                         timeout=float(timeout_s))
        dist.barrier()   # every window is mapped before anyone pushes

    def check_p2p(self):
        """Raises if a peer-memory exchange ever timed out (synchronises the device)."""
        if getattr(self, "_p2p", None) is None:
            return
        import ctypes as C
        from .ceed import b2, lib
        e = C.c_int(0)
        b2(lib.b200_memcpy_d2h(C.byref(e), C.c_void_p(self._p2p["err"]), 4))
        if e.value:
            raise RuntimeError("peer-memory halo exchange timed out waiting for a neighbour")

    def _sum_and_share_p2p(self, Yloc):
        from .ceed import b2, lib
        p = self._p2p
        p["gen"] += 1
        g = p["gen"]
        par = g & 1
        b2(lib.b200_halo_push_signal(p["nn"], p["seg_start"], p["remote"][par], p["rflag"], self.share_cat.data_ptr(),
                                     Yloc.data_ptr(), g))
        b2(lib.b200_halo_wait_unpack(p["nn"], p["flags"], g, self.share_cat.data_ptr(),
                                     p["buf"].value + par * p["total"] * 8, Yloc.data_ptr(), p["total"], p["err"],
                                     p["timeout"]))

    def sum_and_share(self, Yloc):
        """One symmetric exchange replacing ghost->owner ADD followed by owner->ghost INSERT: every rank
        sends its partial sums on ALL shared dofs to every sharer and adds what it receives, so all copies
        end up with the assembled value (allowed harness optimisation)."""
        if getattr(self, "_p2p", None) is not None and Yloc.is_cuda:
            return self._sum_and_share_p2p(Yloc)
        self._exchange(Yloc, self.share_cat, self.share_views, self.share_cat, self.share_views, add=True, tag="sas")

    # ---- split form of sum_and_share: the exchange runs on a side stream while the caller keeps computing on
    # entries that are NOT shared (the interior elements of an operator application)
    def sum_and_share_begin(self, Yloc):
        """Pack the shared dofs (their partial sums must be complete) and start the neighbour exchange."""
        dist = self.dist
        sbuf = self._buffer("ssas", self.share_cat.numel(), Yloc)
        rbuf = self._buffer("rsas", self.share_cat.numel(), Yloc)
        self._gather(sbuf, Yloc, self.share_cat)
        ops = []
        for r in self.neighbours:
            a, b = self.share_views[r]
            ops.append(dist.P2POp(dist.irecv, rbuf[a:b], r))
            ops.append(dist.P2POp(dist.isend, sbuf[a:b], r))
        self._pending = None
        if not ops:
            return
        if Yloc.is_cuda:
            if getattr(self, "_side", None) is None:
                self._side = torch.cuda.Stream(device=Yloc.device)
                self._ev_packed = torch.cuda.Event()
                self._ev_recvd = torch.cuda.Event()
            self._ev_packed.record()                       # pack kernel queued on the compute stream
            self._side.wait_event(self._ev_packed)
            with torch.cuda.stream(self._side):
                works = dist.batch_isend_irecv(ops)
                for w in works:                            # stream-ordered wait: the host does not block
                    w.wait()
                self._ev_recvd.record()
            self._pending = ("cuda", rbuf)
        else:
            self._pending = ("cpu", rbuf, dist.batch_isend_irecv(ops))

    def sum_and_share_end(self, Yloc):
        """Wait for the exchange started by sum_and_share_begin and add the neighbours' partial sums."""
        pend, self._pending = self._pending, None
        if pend is None:
            return
        if pend[0] == "cuda":
            torch.cuda.current_stream(Yloc.device).wait_event(self._ev_recvd)
        else:
            for w in pend[2]:
                w.wait()
        self._scatter(Yloc, self.share_cat, pend[1], True)

    def owner_to_ghost(self, Xloc):
        """DMGlobalToLocal part 2: ghosts receive the owner's value."""
        self._exchange(Xloc, self.own_cat, self.own_views, self.ghost_cat, self.ghost_views, add=False, tag="o2g")

    def ghost_to_owner_add(self, Yloc):
        """DMLocalToGlobal(ADD_VALUES) part 1: owners accumulate the ghosts' partial sums."""
        self._exchange(Yloc, self.ghost_cat, self.ghost_views, self.own_cat, self.own_views, add=True, tag="g2o")
