"""CUDA-graph execution of a whole training step (forward + backward + all-reduce + optimiser).

The step issues several hundred small launches (per-timestep recurrent GEMMs, per-chunk conv GEMMs, loss
kernels); replaying them from one captured graph removes the host launch / ctypes / autograd overhead that would
otherwise dominate a ~ms step (SURVEY.md §7 "Host overhead").  Noise keeps advancing between replays (the torch
CUDA generator is graph-aware) and the Adam bias correction reads a device-resident step counter."""
import torch


def _clone_to_static(batch, device):
    out = {}
    for k, v in batch.items():
        if isinstance(v, dict):
            out[k] = _clone_to_static(v, device)
        elif torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=True).clone()
        else:
            out[k] = v
    return out


def _tensors(batch):
    for v in batch.values():
        if isinstance(v, dict):
            yield from _tensors(v)
        elif torch.is_tensor(v):
            yield v


def _copy_into(static, batch):
    for k, v in batch.items():
        if isinstance(v, dict):
            _copy_into(static[k], v)
        elif torch.is_tensor(v):
            static[k].copy_(v, non_blocking=True)


def _snapshot(objs):
    return [(o, o.snapshot() if hasattr(o, "snapshot") else o.clone()) for o in objs]


@torch.no_grad()
def _restore(snaps):
    for o, s in snaps:
        if hasattr(o, "restore"):
            o.restore(s)
        else:
            o.copy_(s)


class GraphedTrainStep:
    """step_fn(static_batch) must run one full optimisation step and return a scalar device tensor (or None).
    Usage:  g = GraphedTrainStep(step_fn, example_batch); loss = g(next_batch)

    A captured graph freezes everything the HOST decided while it was recorded: Python branches (TACO-RL's
    `current_epoch < bc_epochs` actor loss, `finetune_action_decoder`) and scalars passed by value (`kl_beta`, learning
    rates).  `mode_key` (default: `step_fn.mode_key`) returns a hashable summary of that host state; one graph is kept
    per distinct key and a new one is captured the first time a key is seen, so crossing `bc_epochs` or calling
    `set_kl_beta` takes effect on the next call.  `preserve` (default: `step_fn.preserve`) lists the optimisers /
    tensors a step mutates: they are saved before the warm-up steps a capture needs and restored afterwards, so
    building a graph does not train on the example batch."""

    def __init__(self, step_fn, example_batch, device=None, warmup=3, mode_key=None, preserve=None, buffers=1):
        """buffers=2: two sets of static input buffers, one captured graph per set (sharing one memory pool), used
        alternately: `prefetch()` copies the next batch host->device straight into the set the running step is NOT
        reading, so a host-fed loop needs no staging->static device copy between replays."""
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.statics = [_clone_to_static(example_batch, device) for _ in range(max(1, int(buffers)))]
        self._cur = 0
        self.step_fn = step_fn
        self.warmup = warmup
        self.mode_key = mode_key if mode_key is not None else getattr(step_fn, "mode_key", None)
        self.preserve = preserve if preserve is not None else getattr(step_fn, "preserve", None)
        self._graphs = {}
        self._pool = None
        self.replays = 0
        self.captures = 0
        self._select(self._key())

    @property
    def static(self):
        return self.statics[self._cur]

    def _key(self):
        return self.mode_key() if self.mode_key is not None else None

    def _capture(self, static):
        from . import _lib
        device = self.device
        snaps = _snapshot(self.preserve()) if self.preserve is not None else None
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.step_fn(static)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()      # graphs of one runner never run concurrently: one pool
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph, pool=self._pool):
            out = self.step_fn(static)
            out = out.detach() if torch.is_tensor(out) else None
        # kernels of libtacorl_b200.so recorded in the graph = launched again by every replay
        launches = _lib.launch_count() - n0
        if snaps is not None:
            _restore(snaps)
            torch.cuda.synchronize(device)
        self.captures += 1
        return graph, out, launches

    def _select(self, key):
        k = (key, self._cur)
        if k not in self._graphs:
            self._graphs[k] = self._capture(self.statics[self._cur])
        self.graph, self.out, self.launches_per_replay = self._graphs[k]
        self._active_key = k

    def __call__(self, batch=None):
        """Returns the step's output tensor; with buffers > 1 it is only valid until the next call (shared pool)."""
        main = torch.cuda.current_stream()
        if batch is not None:
            _copy_into(self.static, batch)
        elif self._staged:
            main.wait_event(self._ready)              # the H2D of this step's inputs
            if len(self.statics) > 1:
                self._cur = self._staged_buf           # ... which went straight into the other buffer set
            else:
                _copy_into(self.static, self._staging)
                self._consumed.record(main)
            self._staged = False
        key = self._key()
        if (key, self._cur) != self._active_key:
            self._select(key)
        from . import ops
        ops.refresh_shadows()      # parameters edited through torch since the last replay (checkpoint load, ...)
        self.graph.replay()
        if len(self.statics) > 1 and self._free is not None:
            self._free[self._cur].record(main)         # this buffer set may be refilled once the replay has finished
        self.replays += 1
        return self.out

    _staged = False
    _staging = None
    _free = None

    def prefetch(self, host_batch):
        """Start the host->device copy of the NEXT step's (pinned) batch on a side stream so it overlaps the
        step that is running; the following `__call__()` consumes it (double buffering, like a pinned
        DataLoader with non_blocking copies)."""
        if self._staging is None:
            self._copy_stream = torch.cuda.Stream()
            self._ready = torch.cuda.Event()
            if len(self.statics) > 1:
                self._staging = True
                self._free = [torch.cuda.Event() for _ in self.statics]
                for e in self._free:
                    e.record(torch.cuda.current_stream())
            else:
                self._staging = _clone_to_static(self.static, next(iter(_tensors(self.static))).device)
                self._consumed = torch.cuda.Event()
                self._consumed.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            if len(self.statics) > 1:
                target = (self._cur + 1) % len(self.statics)
                self._copy_stream.wait_event(self._free[target])
                _copy_into(self.statics[target], host_batch)
                self._staged_buf = target
            else:
                self._copy_stream.wait_event(self._consumed)
                _copy_into(self._staging, host_batch)
            self._ready.record(self._copy_stream)
        self._staged = True


def play_lmp_step_fn(module, optimizer, early_step=True):
    """One PlayLMP optimiser step.  early_step: this function runs exactly one backward pass per step(), so the update
    of the parameters behind the encoders may start as soon as their gradients are final (FlatAdam.early_step)."""
    optimizer.early_step = bool(early_step)
    import os
    if os.environ.get("TACORL_EARLY_SUBNET") is not None:      # experiment switches (scripts/early_adam_sweep.sh)
        module.early_subnet_sync = os.environ["TACORL_EARLY_SUBNET"] == "1"
    if os.environ.get("TACORL_EARLY_BG") is not None:
        optimizer.early_background = os.environ["TACORL_EARLY_BG"] == "1"

    def step(batch):
        optimizer.zero_grad(set_to_none=True)
        loss = module.training_step(batch, 0)
        loss.backward()
        optimizer.step()
        return loss
    step.mode_key = lambda: (float(module.kl_beta), float(optimizer.param_groups[0]["lr"]), bool(module.training))
    step.preserve = lambda: [optimizer]
    return step


def tacorl_step_fn(module):
    """Step function of a manual-optimisation module: TACORL or the flat CQL_Offline baseline."""
    def step(batch):
        module.training_step(batch)
        return module.logged.get("train/q1_loss")
    opts = module.optimizers()
    step.mode_key = lambda: (module.current_epoch < module.bc_epochs, bool(getattr(module, "finetune_action_decoder", False)),
                             tuple(float(o.param_groups[0]["lr"]) for o in opts), bool(module.training))
    step.preserve = lambda: list(opts) + [b.flat for b in (module._target_bufs or ())]
    return step
