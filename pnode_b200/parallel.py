"""Batch-sharded data parallelism (one process per GPU, torch.distributed over NCCL / NVLink).

The reference is single-process (`ODEPetsc.comm = PETSc.COMM_SELF`, pnode/petsc_adjoint.py:367); trajectories are
independent, so the forward and adjoint sweeps shard by batch with NO data-path collective.  Only two exchanges exist
(SURVEY.md section 8e):
  1. mu (parameter gradient): one all-reduce(sum) of [np] scalars per backward, because the loss sums over the global batch;
  2. adaptive runs: one scalar all-reduce(sum) of the weighted squared error per step attempt, BEFORE accept/reject, so
     that every rank takes the identical decision and the identical next step (the reference's single shared step size;
     the WRMS norm is over the whole, i.e. global, state vector).
Usage:  ode = ODEPetsc(); ode.comm = BatchComm()   # after torch.distributed.init_process_group(...)
"""
import torch
import torch.distributed as dist


class BatchComm:
    def __init__(self, group=None):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("BatchComm needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.collectives = 0
        self._count_cache = {}
        self.peer = None  # set by enable_peer_reduce()

    def enable_peer_reduce(self):
        """Fuse the mu all-reduce into the tail of the fused adjoint kernels: allocate one symmetric-memory buffer per rank
        (torch symmetric memory = CUDA VMM allocations mapped into every peer over NVLink), exchange the mappings once,
        and hand the peer address table to the kernels (include/pnode_b200.h, *_adjoint_dp).  Returns True when active."""
        if self.world == 1 or self.peer is not None:
            return self.peer is not None
        from . import _lib
        import torch.distributed._symmetric_memory as symm_mem

        nbytes = int(_lib.load().pnode_peer_buffer_bytes(self.world))
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        buf.zero_()
        handle = symm_mem.rendezvous(buf, dist.group.WORLD if self.group is None else self.group)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.peer = {"buf": buf, "handle": handle, "ptrs_dev": int(handle.buffer_ptrs_dev), "epoch": 0}
        return True

    def next_epoch(self):
        self.peer["epoch"] += 1
        self.collectives += 1
        return self.peer["epoch"]

    def reserve_epochs(self, n):
        """First of n consecutive collective numbers for kernels that exchange through the symmetric inbox themselves (the
        conv block's statistics exchange, csrc/conv_block.cu); every rank reserves the same numbers in the same order."""
        first = self.peer["epoch"] + 1
        self.peer["epoch"] += n
        self.collectives += n
        return first

    def allreduce_sum(self, tensor):
        if self.world > 1:
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
            self.collectives += 1
        return tensor

    _comm_stream = None

    def allreduce_sum_layered(self, mu, cb, offset):
        """All-reduce of mu where the part owned by the evaluator `cb` (mu[offset:], csrc/dense_mlp.cu) goes out LAYER BY
        LAYER on a side stream, each slice as soon as the event of its last contribution has fired -- the sweep's last
        vector-Jacobian product differentiates the layers last to first, so the collective of layer l runs while layers
        l-1 .. 0 are still being computed (BASELINE config 5: 298 MB of mu, five slices).  The rest of mu follows on the
        caller's stream, which then waits for the side stream."""
        slices = cb.layer_slices(mu[offset:]) if (self.world > 1 and hasattr(cb, "layer_slices")) else None
        if not slices:
            return self.allreduce_sum(mu)
        main = torch.cuda.current_stream()
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        cs = self._comm_stream
        covered = []
        with torch.cuda.stream(cs):
            for ev, lo, n in slices:
                cs.wait_event(ev)
                dist.all_reduce(mu[offset + lo: offset + lo + n], op=dist.ReduceOp.SUM, group=self.group)
                self.collectives += 1
                covered.append((offset + lo, offset + lo + n))
        pos = 0
        for lo, hi in sorted(covered) + [(mu.numel(), mu.numel())]:
            if lo > pos:
                dist.all_reduce(mu[pos:lo], op=dist.ReduceOp.SUM, group=self.group)
                self.collectives += 1
            pos = max(pos, hi)
        main.wait_stream(cs)
        return mu

    def allreduce_scalar(self, tensor):
        return self.allreduce_sum(tensor)

    def same_on_all_ranks(self, n_local):
        """True when every rank holds the same count (one MAX all-reduce of (n, -n), remembered per n)."""
        if self.world == 1:
            return True
        key = ("same", n_local)
        if key not in self._count_cache:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else "cpu"
            c = torch.tensor([n_local, -n_local], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.MAX, group=self.group)
            self._count_cache[key] = int(c[0].item()) == -int(c[1].item())
        return self._count_cache[key]

    def global_count(self, n_local):
        """Global state length N of the WRMS norm (ranks may hold ragged shards)."""
        if self.world == 1:
            return n_local
        if n_local not in self._count_cache:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else "cpu"
            c = torch.tensor([n_local], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=self.group)
            self._count_cache[n_local] = int(c.item())
        return self._count_cache[n_local]


def shard_batch(tensor, rank, world, dim=0):
    """Contiguous batch slice of `tensor` owned by `rank` (ragged when world does not divide the batch)."""
    n = tensor.shape[dim]
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    length = base + (1 if rank < rem else 0)
    return tensor.narrow(dim, start, length)
