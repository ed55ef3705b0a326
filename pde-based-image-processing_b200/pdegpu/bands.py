"""One very large image on several GPUs: column bands with a per-sweep halo exchange (SURVEY.md 8e,
BASELINE.json configs[4]: "one 16384x16384 image (row-band halo exchange) at 1/2/4/8 B200").

The image is cut along the slow axis (Matlab column index j) into `world` contiguous bands whose first
columns are EVEN (so a pixel's red/black colour is the same in band and image coordinates). A rank keeps
its band plus H = 2*T halo columns per inner side, for every field. One exchange step = each rank sends
its outermost H owned columns of every UNKNOWN to the neighbour's halo (coefficients are static: scattered
once, halo included). After an exchange the rank runs T red-black sweeps on its local array as if it were
a whole image; a sweep spoils at most 2 columns from each cut (the local "border" column and the black
pixels next to it), so after T sweeps exactly the H halo columns are stale and the owned columns equal,
bit for bit, what a single GPU computes. Larger T = fewer, larger messages (temporal blocking of the halo).

Only solver 1 (point red-black) splits this way; lines of the line solver along j cross the cuts.
torch.distributed is plumbing (NCCL send/recv of contiguous column blocks over NVLink on the GPU box,
gloo in the CPU tests); the sweep itself is injected, so the host logic is testable without a GPU.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple


def band_columns(ncols: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous column ranges [j0, j1) with even j0, as equal as that allows."""
    if world < 1 or ncols < 2 * world:
        raise ValueError("need at least 2 columns per band")
    cuts = [0]
    for k in range(1, world):
        c = (ncols * k // world) & ~1
        cuts.append(max(c, cuts[-1] + 2))
    cuts.append(ncols)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


class BandPlan:
    """Geometry of one rank's band: owned columns [j0, j1), local array columns [a0, a1) incl. halos."""

    def __init__(self, nrows: int, ncols: int, rank: int, world: int, sweeps_per_exchange: int = 1):
        self.nrows, self.ncols, self.rank, self.world = nrows, ncols, rank, world
        self.T = int(sweeps_per_exchange)
        self.H = 2 * self.T
        self.j0, self.j1 = band_columns(ncols, world)[rank]
        bands = band_columns(ncols, world)
        if world > 1 and min(b - a for a, b in bands) < self.H:
            raise ValueError("bands narrower than the halo")
        self.left = rank - 1 if rank > 0 else None
        self.right = rank + 1 if rank < world - 1 else None
        self.a0 = self.j0 - (self.H if self.left is not None else 0)
        self.a1 = self.j1 + (self.H if self.right is not None else 0)
        self.local_cols = self.a1 - self.a0
        self.own = slice(self.j0 - self.a0, self.j1 - self.a0)        # owned columns inside the local array

    def take_local(self, full):
        """Local block (halo included) of a full field given as [ncols, nrows] (column-major image)."""
        return full[self.a0:self.a1]


def exchange_halos(plan: BandPlan, unknowns: Sequence, group=None) -> int:
    """Send the outermost H owned columns of every unknown to the neighbours, receive theirs into the halos.
    `unknowns`: tensors [local_cols, nrows] (contiguous). Returns the number of bytes this rank sent."""
    import torch.distributed as dist
    if plan.world == 1:
        return 0
    H, o = plan.H, plan.own
    ops, sent = [], 0
    for x in unknowns:
        if plan.left is not None:
            ops.append(dist.P2POp(dist.isend, x[o.start:o.start + H], plan.left, group))
            ops.append(dist.P2POp(dist.irecv, x[o.start - H:o.start], plan.left, group))
            sent += H * x.shape[1] * x.element_size()
        if plan.right is not None:
            ops.append(dist.P2POp(dist.isend, x[o.stop - H:o.stop], plan.right, group))
            ops.append(dist.P2POp(dist.irecv, x[o.stop:o.stop + H], plan.right, group))
            sent += H * x.shape[1] * x.element_size()
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    return sent


def relax_bands(plan: BandPlan, unknowns: Sequence, sweep: Callable[[int], None], iters: int, group=None,
                before_sweep: Callable[[], None] | None = None) -> int:
    """`iters` red-black sweeps of the band: exchange, then up to T sweeps, until done. `sweep(n)` runs n sweeps
    on the rank's local arrays (libpdegpu: pdegpu_dev_relax(..., iter=n, solver=1)). `before_sweep` (optional)
    is called after every exchange, e.g. to make the sweep's stream wait for the communication stream."""
    done, sent = 0, 0
    while done < iters:
        n = min(plan.T, iters - done)
        sent += exchange_halos(plan, unknowns, group)
        if before_sweep:
            before_sweep()
        sweep(n)
        done += n
    return sent


class GpuBand:
    """A rank's band of one flow / PDE problem in device memory, relaxed with libpdegpu's point solver.
    fields: dict name -> torch tensor [local_cols, nrows] on this rank's GPU (halo columns included);
    unknown names first in `unknowns`."""

    def __init__(self, ctx, plan: BandPlan, family: int, fields: Dict[str, "object"], group=None, transport: str = "nccl"):
        """transport "nccl": halo columns through torch.distributed send / recv (exchange_halos); "p2p": libpdegpu's own
        exchange (pdegpu_band_*: peer stores + flags over NVLink, no host synchronisation per step), set up with
        connect_p2p() (one process per GPU) or connect_local() (several bands in one process)."""
        import torch
        from . import lib
        self.ctx, self.plan, self.family, self.f, self.group = ctx, plan, family, fields, group
        self.transport = transport
        self.xchg = None
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=next(iter(fields.values())).device)
        nr, lc = plan.nrows, plan.local_cols
        p = lambda k: fields[k].data_ptr()
        if family == lib.FLOW_LLIN4:
            self.unknowns = [fields["dU"], fields["dV"]]
            self.sys = lib.make_system(family, nr, lc, x=(p("dU"), p("dV")), x0=(p("U"), p("V")), m=p("M"),
                                       c=(p("Cu"), p("Cv")), d=(p("Du"), p("Dv")), w=[p(k) for k in ("wW", "wN", "wE", "wS")])
        elif family == lib.FLOW_ELIN4:
            self.unknowns = [fields["U"], fields["V"]]
            self.sys = lib.make_system(family, nr, lc, x=(p("U"), p("V")), m=p("M"),
                                       c=(p("Cu"), p("Cv")), d=(p("Du"), p("Dv")), w=[p(k) for k in ("wW", "wN", "wE", "wS")])
        elif family == lib.PDE4:
            self.unknowns = [fields["X"]]
            self.sys = lib.make_system(family, nr, lc, x=(p("X"),), c=(p("B"),), d=(p("TRACE"),), w=[p(k) for k in ("wW", "wN", "wE", "wS")])
        else:
            raise ValueError("band split is built for the flow and PDE4 families")

    def _make_exchange(self):
        from . import lib
        p = self.plan
        if self.xchg is None:
            self.xchg = lib.BandExchange(self.ctx, p.nrows, p.H, len(self.unknowns), p.left is not None, p.right is not None)
        return self.xchg

    def connect_p2p(self):
        """one process per GPU: all-gather the mailboxes' IPC handles over torch.distributed, open the neighbours'"""
        import torch
        import torch.distributed as dist
        x = self._make_exchange()
        mine = torch.frombuffer(bytearray(x.export()), dtype=torch.uint8).to(self.unknowns[0].device)
        everyone = [torch.empty_like(mine) for _ in range(self.plan.world)]
        dist.all_gather(everyone, mine, group=self.group)
        h = lambda r: bytes(everyone[r].cpu().numpy().tobytes()) if r is not None else None
        x.connect(h(self.plan.left), h(self.plan.right))
        dist.barrier(group=self.group)

    def connect_local(self, left: "GpuBand | None", right: "GpuBand | None"):
        self._make_exchange().connect_local(left._make_exchange() if left is not None else None,
                                            right._make_exchange() if right is not None else None)

    def exchange_p2p(self):
        p = self.plan
        self.xchg.exchange([u.data_ptr() for u in self.unknowns], p.own.start, p.own.stop)

    def relax(self, iters: int, omega: float) -> int:
        """`iters` red-black sweeps with halo exchanges, everything ordered through the context's stream."""
        import torch
        if self.transport == "p2p":
            done, before = 0, self.xchg.bytes_sent if self.xchg else 0
            while done < iters:
                n = min(self.plan.T, iters - done)
                if self.plan.world > 1:
                    self.exchange_p2p()
                self.ctx.relax(self.sys, n, omega, 1)
                done += n
            return (self.xchg.bytes_sent - before) if self.xchg else 0
        with torch.cuda.stream(self.stream):
            return relax_bands(self.plan, self.unknowns, lambda n: self.ctx.relax(self.sys, n, omega, 1), iters, self.group)
