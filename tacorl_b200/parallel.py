"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce (sum) of the flat gradient
buffer over NVLink/NVSwitch, 1/world folded into the fused Adam kernel.  Replaces Lightning's DDP-over-gloo
(/root/reference/config/trainer/default.yaml:1-3, scripts/train.py:73-75).  Works on CPU tensors with the
gloo backend (tests/test_parallel_gloo.py)."""
import os

import torch
import torch.distributed as dist

SM_RESERVE_FOR_COLLECTIVES = 16   # SMs left out of the persistent conv grids when gradients are exchanged (one per NCCL channel)


def collective_channels(world_size):
    """NCCL channels = SMs reserved for the overlapped gradient all-reduce.  Measured (ms/step, PlayLMP bf16, 64 windows
    per GPU): N = 2 (round 1, scripts/n2_reserve_sweep.sh): 8 -> 4.30, 16 -> 3.96, 24 -> 4.01, 32 -> 4.07;
    N = 8 (round 2, scripts/n8_sweep.sh): 8 -> 3.915, 16 -> 3.905, 24 -> 3.79, 32 -> 3.79 (bf16 wire 3.77, tree 4.65)."""
    forced = os.environ.get("TACORL_NCCL_CHANNELS")
    if forced:
        return int(forced)
    return 16 if world_size <= 2 else 24


class BucketedAllReduce:
    """All-reduces a flat gradient buffer in fixed-size buckets on a side stream (CUDA) so the exchange of
    early buckets overlaps whatever the caller still runs on the main stream.
    wire_dtype=torch.bfloat16: every bucket is cast to bf16, summed over the ranks in bf16 and cast back into the fp32
    buffer (half the NVLink bytes; used with the bf16 compute path, whose gradients carry bf16 operand rounding anyway)."""

    def __init__(self, world_size, bucket_elems=16 << 20, group=None, wire_dtype=None):
        self.world = world_size
        self.bucket = bucket_elems
        self.group = group
        self.wire_dtype = wire_dtype
        self._stream = None
        self._stage = None

    def _reduce(self, flat):
        for o in range(0, flat.numel(), self.bucket):
            chunk = flat[o:o + self.bucket]
            if self.wire_dtype is not None and chunk.is_cuda:
                if self._stage is None or self._stage.numel() < self.bucket or self._stage.device != chunk.device:
                    self._stage = torch.empty(self.bucket, device=chunk.device, dtype=self.wire_dtype)
                st = self._stage[:chunk.numel()]
                st.copy_(chunk)
                dist.all_reduce(st, op=dist.ReduceOp.SUM, group=self.group)
                chunk.copy_(st)
            else:
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)

    def start(self, flat):
        """Enqueue the all-reduce of `flat` on the side stream (after everything already queued on the current
        stream) and return immediately; `finish()` makes the current stream wait for it."""
        if self.world <= 1 or flat.numel() == 0:
            return
        if flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=flat.device)
            self._device = flat.device
            self._stream.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self._stream):
                self._reduce(flat)
        else:
            self._reduce(flat)

    def finish(self):
        if self._stream is not None:
            torch.cuda.current_stream(self._device).wait_stream(self._stream)

    def wait_on(self, stream):
        """`stream` waits for every exchange enqueued so far."""
        if self._stream is not None:
            stream.wait_stream(self._stream)

    _device = None

    def __call__(self, flat):
        self.start(flat)
        self.finish()


class NativeAllReduce:
    """The same exchange through the library's own NCCL communicator (tacorl_dp_allreduce_{init,enqueue,wait}): no
    torch.distributed call on the data path.  torch.distributed (any backend) is only used once, to hand rank 0's
    ncclUniqueId to the other ranks."""
    _ready = False

    @classmethod
    def ensure_init(cls, world_size, group=None):
        if cls._ready:
            return
        import ctypes
        from . import _lib
        rank = dist.get_rank(group)
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _lib.call("tacorl_dp_unique_id", buf)
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        _lib.call("tacorl_dp_allreduce_init", ctypes.create_string_buffer(box[0], 128), rank, world_size)
        cls._ready = True

    @classmethod
    def shutdown(cls):
        if cls._ready:
            from . import _lib
            _lib.call("tacorl_dp_allreduce_destroy")
            cls._ready = False

    def __init__(self, world_size, bucket_elems=16 << 20, group=None):
        self.world, self.bucket, self.group = world_size, bucket_elems, group
        self.ensure_init(world_size, group)

    def start(self, flat):
        if self.world <= 1 or flat.numel() == 0:
            return
        from . import _lib
        assert flat.is_cuda and flat.dtype in (torch.float32, torch.bfloat16)
        for o in range(0, flat.numel(), self.bucket):
            chunk = flat[o:o + self.bucket]
            _lib.call("tacorl_dp_allreduce_enqueue", _lib.ptr_any(chunk), chunk.numel(),
                      1 if chunk.dtype == torch.bfloat16 else 0, _lib.stream())

    def wait_on(self, stream):
        """`stream` waits for every exchange enqueued so far."""
        from . import _lib
        with torch.cuda.stream(stream):
            _lib.call("tacorl_dp_allreduce_wait", _lib.stream())

    def finish(self):
        from . import _lib
        _lib.call("tacorl_dp_allreduce_wait", _lib.stream())

    def __call__(self, flat):
        self.start(flat)
        self.finish()


def attach_data_parallel(optimizer, world_size, group=None, bucket_elems=16 << 20, wire_dtype="auto", native="auto"):
    """Make a FlatAdam average its gradient over `world_size` ranks before the update (DDP semantics:
    mean of local-mean gradients; clip-by-norm acts on the reduced gradient, cql_offline_lightning.py:521-537).
    wire_dtype: "auto" = fp32 unless TACORL_WIRE=bf16; torch.bfloat16 = compressed exchange; None = fp32.
    native: exchange through the library's own NCCL communicator (tacorl_dp_allreduce_*; CUDA, fp32 wire) instead of
    torch.distributed's; "auto" = yes unless TACORL_NATIVE_NCCL=0."""
    if wire_dtype == "auto":
        # measured (profiles/r02): at N = 2 the two cast passes cost more than the halved NVLink bytes save (3.50 -> 3.60
        # ms/step), so fp32 stays the default; TACORL_WIRE=bf16 selects the compressed exchange
        flat0 = getattr(optimizer, "flat_params", None)
        want = os.environ.get("TACORL_WIRE", "fp32") == "bf16"
        wire_dtype = torch.bfloat16 if (want and flat0 is not None and flat0.is_cuda) else None
    flat1 = getattr(optimizer, "flat_params", None)
    if native == "auto":
        native = os.environ.get("TACORL_NATIVE_NCCL", "1") != "0"
    if native and world_size > 1 and flat1 is not None and flat1.is_cuda and wire_dtype is None:
        optimizer.grad_sync = NativeAllReduce(world_size, bucket_elems, group)
    else:
        optimizer.grad_sync = BucketedAllReduce(world_size, bucket_elems, group, wire_dtype)
    optimizer.grad_scale = 1.0 / world_size
    flat = getattr(optimizer, "flat_params", None)
    if world_size > 1 and flat is not None and flat.is_cuda and "TACORL_SM_RESERVE" not in os.environ:
        # the all-reduce of the non-encoder slice overlaps the encoder backward: the persistent convolution kernels
        # launched between the start of that exchange and step() leave one SM per NCCL channel free (FlatAdam sets and
        # clears the reserve around that window; the forward pass and everything before the exchange use all SMs)
        optimizer.sm_reserve = collective_channels(world_size)
    return optimizer


def shard_batch(batch, rank, world_size):
    """Contiguous split of a replay batch over ranks."""
    def sl(t):
        n = t.shape[0]
        assert n % world_size == 0, f"batch of {n} windows does not split evenly over {world_size} ranks"
        per = n // world_size
        return t[rank * per:(rank + 1) * per]
    return {k: ({kk: sl(vv) for kk, vv in v.items()} if isinstance(v, dict) else sl(v)) for k, v in batch.items()}
