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

    def __init__(self, step_fn, example_batch, device=None, warmup=3, mode_key=None, preserve=None):
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.static = _clone_to_static(example_batch, device)
        self.step_fn = step_fn
        self.warmup = warmup
        self.mode_key = mode_key if mode_key is not None else getattr(step_fn, "mode_key", None)
        self.preserve = preserve if preserve is not None else getattr(step_fn, "preserve", None)
        self._graphs = {}
        self.replays = 0
        self.captures = 0
        self._select(self._key())

    def _key(self):
        return self.mode_key() if self.mode_key is not None else None

    def _capture(self):
        from . import _lib
        device = self.device
        snaps = _snapshot(self.preserve()) if self.preserve is not None else None
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.step_fn(self.static)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            out = self.step_fn(self.static)
            out = out.detach() if torch.is_tensor(out) else None
        # kernels of libtacorl_b200.so recorded in the graph = launched again by every replay
        launches = _lib.launch_count() - n0
        if snaps is not None:
            _restore(snaps)
            torch.cuda.synchronize(device)
        self.captures += 1
        return graph, out, launches

    def _select(self, key):
        if key not in self._graphs:
            self._graphs[key] = self._capture()
        self.graph, self.out, self.launches_per_replay = self._graphs[key]
        self._active_key = key

    def __call__(self, batch=None):
        key = self._key()
        if key != self._active_key:
            self._select(key)
        if batch is not None:
            _copy_into(self.static, batch)
        elif self._staged:
            # inputs prefetched by `prefetch()`: wait for the H2D, move them into the graph's static buffers
            main = torch.cuda.current_stream()
            main.wait_event(self._ready)
            _copy_into(self.static, self._staging)
            self._consumed.record(main)
            self._staged = False
        from . import ops
        ops.refresh_shadows()      # parameters edited through torch since the last replay (checkpoint load, ...)
        self.graph.replay()
        self.replays += 1
        return self.out

    _staged = False
    _staging = None

    def prefetch(self, host_batch):
        """Start the host->device copy of the NEXT step's (pinned) batch on a side stream so it overlaps the
        step that is running; the following `__call__()` consumes it (double buffering, like a pinned
        DataLoader with non_blocking copies)."""
        if self._staging is None:
            self._staging = _clone_to_static(self.static, next(iter(_tensors(self.static))).device)
            self._copy_stream = torch.cuda.Stream()
            self._ready = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed)
            _copy_into(self._staging, host_batch)
            self._ready.record(self._copy_stream)
        self._staged = True


def play_lmp_step_fn(module, optimizer, early_step=True):
    """One PlayLMP optimiser step.  early_step: this function runs exactly one backward pass per step(), so the update
    of the parameters behind the encoders may start as soon as their gradients are final (FlatAdam.early_step)."""
    optimizer.early_step = bool(early_step)

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
    def step(batch):
        module.training_step(batch)
        return module.logged.get("train/q1_loss")
    opts = module.optimizers()
    step.mode_key = lambda: (module.current_epoch < module.bc_epochs, bool(module.finetune_action_decoder),
                             tuple(float(o.param_groups[0]["lr"]) for o in opts), bool(module.training))
    step.preserve = lambda: list(opts) + [b.flat for b in (module._target_bufs or ())]
    return step
