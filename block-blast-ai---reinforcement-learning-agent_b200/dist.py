"""Multi-GPU plumbing: one process per GPU (torchrun), NCCL over NVLink for the only real
exchange on this path — the PPO gradient all-reduce — plus env-shard bookkeeping.

Env instances shard embarrassingly: rank r owns global env ids
[r*n_local, (r+1)*n_local); trajectories depend on (seed, global id, actions) only, so no
collective touches the env path.  Collectives: (1) one flat-bucket all-reduce of the
5,290,113 fp32 gradients per optimiser step, (2) a 3-double all-reduce for the whole-buffer
advantage normalisation (rollout.py), (3) a few scalars for metrics.  The reference is
single-process, so it has no counterpart for this file.
"""
import os

import torch
import torch.distributed as td


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment; no-op for world size 1.
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            td.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            td.init_process_group(backend)
    return rank, world, local


def world_size():
    return td.get_world_size() if td.is_available() and td.is_initialized() else 1


def rank():
    return td.get_rank() if td.is_available() and td.is_initialized() else 0


def shard(total_envs, rank_=None, world=None):
    """(global_env_offset, n_local) of this rank's env shard; the remainder goes to the low ranks."""
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    base, rem = divmod(int(total_envs), w)
    n_local = base + (1 if r < rem else 0)
    offset = r * base + min(r, rem)
    return offset, n_local


class FlatGradBucket:
    """All parameters' gradients as views into ONE contiguous buffer, so the per-step gradient
    exchange needs no packing copies: one NCCL all-reduce of the whole buffer (21.2 MB fp32), or —
    with ``early_from`` — two: parameters ``early_from:`` (the FC encoder and the heads: 4.7 M of the
    5.29 M values, whose gradients are complete as soon as backward leaves the FC layers) are reduced
    asynchronously on NCCL's stream while backward is still inside the convolution stack (most of the
    step's compute); the convolution slice follows when backward ends."""

    def __init__(self, params, early_from=None):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self._view(p, off)
            off += p.numel()
        self.split = None
        self._work, self._pending, self._armed = None, 0, False
        if early_from is not None and 0 < early_from < len(self.params):
            self.split = sum(p.numel() for p in self.params[:early_from])
            self._early = self.params[early_from:]
            for p in self._early:
                p.register_post_accumulate_grad_hook(self._on_grad)

    def _view(self, p, off):
        # same sizes AND strides as the parameter (channels_last conv weights keep their
        # layout, which is what autograd's gradient layout contract wants)
        return self.flat[off:off + p.numel()].as_strided(p.size(), p.stride())

    def _on_grad(self, _p):
        if not self._armed:
            return
        self._pending -= 1
        if self._pending == 0 and world_size() > 1:
            # every gradient of the early slice is final: start its reduction behind the work queued so far
            self._work = td.all_reduce(self.flat[self.split:], async_op=True)

    def zero(self):
        """Zero the gradients; also arms the early-slice reduction for the backward that follows."""
        self.flat.zero_()
        if self.split is not None:
            self._pending, self._armed, self._work = len(self._early), True, None

    def rebind(self):
        """Re-attach the views if something replaced p.grad (e.g. zero_grad(set_to_none=True))."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat[off:off + 1].data_ptr() or p.grad.stride() != p.stride():
                g = self._view(p, off)
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g
            off += p.numel()

    def all_reduce_mean(self):
        w = world_size()
        self._armed = False
        if w > 1:
            if self._work is not None:
                td.all_reduce(self.flat[:self.split])
                self._work.wait()
                self._work = None
            else:
                td.all_reduce(self.flat)
            self.flat.div_(w)

    def clip_grad_norm_(self, max_norm):
        """torch.nn.utils.clip_grad_norm_ (L2) over all parameters, on the flat buffer: one norm and
        one scale kernel instead of a foreach over every tensor.  Returns the total norm."""
        total = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        return total


def broadcast_module(module, src=0):
    """Rank `src`'s parameters and buffers to everyone (start of training)."""
    if world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            td.broadcast(t.data, src)


def all_reduce_scalars(values, op="sum", device=None):
    """All-reduce a short list of Python numbers; returns a list of floats."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64,
                     device=device or ("cuda" if torch.cuda.is_available() and td.is_initialized() and td.get_backend() == "nccl" else "cpu"))
    if world_size() > 1:
        td.all_reduce(t, op={"sum": td.ReduceOp.SUM, "max": td.ReduceOp.MAX, "min": td.ReduceOp.MIN}[op])
    return t.tolist()


def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def pin_to_gpu_numa(local_rank, local_world):
    """Bind this process to host cores next to its GPU: the cores of the GPU's NUMA node (from sysfs via
    the PCI address), split evenly among the ranks whose GPUs share that node.  Pinned host buffers
    allocated afterwards are first-touched, hence placed, on that node.  The host-buffer step
    (bb_env_step_host) is one PCIe transfer + a stream synchronise per step per rank; without the
    binding eight ranks' Python threads and copy completions migrate over all cores of the box.
    Returns a dict describing what was done (for the bench line); never raises."""
    info = {"numa_node": None, "cpus": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        n_gpus = pynvml.nvmlDeviceGetCount()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        order = [int(x) for x in visible.split(",")] if visible and all(x.strip().isdigit() for x in visible.split(",")) else list(range(n_gpus))

        def node_of(idx):
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(order[idx])).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            bus = bus.lower()
            if len(bus.split(":")[0]) == 8:            # nvml prints an 8-digit domain, sysfs a 4-digit one
                bus = bus[4:]
            with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
                return int(f.read().strip())

        nodes = [node_of(i) for i in range(local_world)]
        node = nodes[local_rank]
        path = "/sys/devices/system/node/node%d/cpulist" % max(node, 0)
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in _parse_cpulist(open(path).read()) if c in allowed] if os.path.exists(path) else allowed
        if not cpus:
            cpus = allowed
        peers = [r for r in range(local_world) if nodes[r] == node]
        k = peers.index(local_rank)
        per = max(1, len(cpus) // len(peers))
        mine = cpus[k * per:(k + 1) * per] or cpus
        os.sched_setaffinity(0, mine)
        info = {"numa_node": node, "cpus": "%d-%d (%d cores)" % (mine[0], mine[-1], len(mine)), "gpu_numa_nodes": nodes}
    except Exception as e:                            # no nvml / sysfs: leave the affinity alone
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info
