"""Bucketed gradient all-reduce overlapped with the backward pass."""
import os

import torch
import torch.distributed as dist


def init_process_group(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment)"""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        if os.environ.get("DENET_NUMA_BIND", "1") != "0":
            bind_to_gpu_numa_node(local)
        dist.init_process_group(backend, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def _pci_bus_id(local_rank):
    """'0000:1b:00.0'-style PCI address of a CUDA device through the runtime API"""
    import ctypes
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            rt = None
    if rt is None:
        return None
    buf = ctypes.create_string_buffer(32)
    if rt.cudaDeviceGetPCIBusId(buf, 32, int(local_rank)) != 0:
        return None
    return buf.value.decode().lower()


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and therefore its pinned-host allocations, first touch) to the CPUs of the NUMA node its GPU
    is attached to.  Round 1 ran all 8 ranks of a box on NUMA node 0: four GPUs pulled their 100 MB image batches
    across the socket interconnect, and end-to-end scaling fell to 0.82 at N = 8.  Returns the node or None."""
    try:
        bus = _pci_bus_id(local_rank)
        if bus is None:
            return None
        node_path = "/sys/bus/pci/devices/%s/numa_node" % bus
        if not os.path.exists(node_path):
            return None
        node = int(open(node_path).read().strip())
        if node < 0:
            return None
        cpulist = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = (cpus & allowed) or cpus          # the launcher may have restricted us to the other node: override
        os.sched_setaffinity(0, use)
        return node
    except (OSError, ValueError):
        return None


def shard_batch(global_count, rank, world):
    """images [rank*B, (rank+1)*B) of the global batch go to rank `rank` (the reference hands batch `index` to worker
    index % n round-robin, model/train_multi.py:113-119)"""
    per = global_count // world
    return rank * per, (rank + 1) * per


def broadcast_state(tensors, src=0, group=None):
    """make every rank's copy of `tensors` equal to rank `src`'s: one flat broadcast per dtype (the tensors are small
    and many: per-tensor broadcasts would be ~hundreds of collectives).  Returns the number of tensors synchronised
    (0 when torch.distributed is not initialised or the world is a single rank)."""
    if not dist.is_initialized() or dist.get_world_size(group) <= 1 or not tensors:
        return 0
    by_kind = {}
    for t in tensors:
        by_kind.setdefault((t.dtype, t.device), []).append(t)
    with torch.no_grad():
        for (_, _), ts in by_kind.items():
            flat = torch.cat([t.detach().reshape(-1) for t in ts])
            dist.broadcast(flat, src=src, group=group)
            o = 0
            for t in ts:
                n = t.numel()
                t.detach().copy_(flat[o:o + n].view(t.shape))
                o += n
    return len(tensors)


def state_checksum(tensors):
    """order-sensitive fp64 checksum of a tensor list (used to assert that replicas hold the same state)"""
    acc = 0.0
    for i, t in enumerate(tensors):
        acc += float(t.detach().double().sum().item()) * (1.0 + 1e-3 * (i % 97))
    return acc


def plan_buckets(layer_ranges, bucket_elems):
    """layer_ranges: [(layer_index, start, end)] element ranges of a flat gradient buffer, ascending in layer order.
    Gradients become final in DESCENDING layer order during backward; returns [(ready_after_layer, start, end)]:
    bucket i may be reduced once backward has finished layer `ready_after_layer`."""
    buckets = []
    cur_end = None
    cur_start = None
    for layer_index, start, end in sorted(layer_ranges, key=lambda r: -r[1]):
        if end <= start:
            continue
        if cur_end is None:
            cur_end = end
        cur_start = start
        if cur_end - cur_start >= bucket_elems:
            buckets.append((layer_index, cur_start, cur_end))
            cur_end = None
    if cur_end is not None:
        buckets.append((min(r[0] for r in layer_ranges), cur_start, cur_end))
    return buckets


class GradientAllReduce:
    """all-reduce(sum) of a flat fp32 gradient buffer in buckets, launched on a side stream as the backward pass
    retires the layers that own each bucket; the solver divides by the world size (grad_scale)."""

    def __init__(self, flat_grad, layer_ranges, bucket_bytes=32 << 20, extra_mean_tensors=None, group=None):
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = plan_buckets(layer_ranges, max(1, bucket_bytes // 4))
        self.extra = extra_mean_tensors or []     # e.g. batch-norm running statistics: averaged once per step
        self.cuda = flat_grad.is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        self._next = 0
        self._works = []
        self._extra_flat = None

    def begin_step(self):
        self._next = 0
        self._works = []

    def _launch(self, start, end):
        view = self.flat[start:end]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self._works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def will_launch(self, layer_index):
        """True when layer_done(layer_index) is going to start the all-reduce of at least one bucket (the caller must
        have finished every pending gradient write of the layers >= layer_index before that)"""
        return self.world > 1 and self._next < len(self.buckets) and self.buckets[self._next][0] >= layer_index

    def layer_done(self, layer_index):
        """called by the backward pass after layer `layer_index` has written its gradients"""
        if self.world <= 1:
            return
        while self._next < len(self.buckets) and self.buckets[self._next][0] >= layer_index:
            _, start, end = self.buckets[self._next]
            self._launch(start, end)
            self._next += 1

    def finish_step(self):
        """flush the remaining buckets, average the extra tensors, make the compute stream wait; returns the factor
        the solver applies to the summed gradients"""
        if self.world <= 1:
            return 1.0
        while self._next < len(self.buckets):
            _, start, end = self.buckets[self._next]
            self._launch(start, end)
            self._next += 1
        if self.extra:
            # one contiguous tensor (ModelCNN keeps the running statistics of all batch-norm layers in one flat buffer):
            # reduce and scale it in place - no gather / scatter kernels per tensor
            inplace = len(self.extra) == 1 and self.extra[0].is_contiguous()
            flat = self.extra[0].view(-1) if inplace else torch.cat([t.reshape(-1) for t in self.extra])
            if self.cuda:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                with torch.cuda.stream(self.stream):
                    self.stream.wait_event(ev)
                    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        else:
            for w in self._works:
                w.wait()
        if self.extra:
            flat.mul_(1.0 / self.world)
            if not inplace:
                o = 0
                for t in self.extra:
                    t.copy_(flat[o:o + t.numel()].view(t.shape))
                    o += t.numel()
        return 1.0 / self.world
