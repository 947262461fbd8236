"""Launch-side drop-in for the part of `pl.Trainer` the reference's scripts/train.py:28-75 uses
(/root/reference/config/trainer/default.yaml:1-4: `strategy: ddp`, all GPUs, `precision: 16`; rl.yaml:4 `max_steps`):
one process per GPU (torchrun), NCCL gradient exchange through tacorl_b200.parallel instead of Lightning's DDP over
gloo, the whole optimisation step replayed from a CUDA graph, and checkpoints in Lightning's on-disk format (`epoch`,
`global_step`, `state_dict`, `optimizer_states`, `hyper_parameters`), so that runs resume from / are resumed by the
reference (`trainer.fit(model, ckpt_path=last.ckpt)`, train.py:48-66; utils/networks.py:90-117).

`B200Strategy` is the object to hand to a real `pl.Trainer(strategy=...)` when pytorch_lightning is installed; it is
only defined then.  This image has no pytorch_lightning, so `Trainer` below is what the tests and bench.py exercise."""
import os
from pathlib import Path

import torch
import torch.distributed as dist

from . import ops, parallel, runtime

CKPT_VERSION = "1.6.5"      # the reference pins pytorch-lightning < 1.7 (setup.cfg:17)


def _to_device(batch, device):
    return {k: (_to_device(v, device) if isinstance(v, dict) else (v.to(device) if torch.is_tensor(v) else v))
            for k, v in batch.items()}


def _optimizers_of(module):
    opts = module.optimizers() if hasattr(module, "optimizers") else module.configure_optimizers()
    return list(opts) if isinstance(opts, (list, tuple)) else [opts]


def save_checkpoint(module, path, epoch=0, global_step=0, extra=None):
    """Lightning-format checkpoint of a tacorl_b200 module (FlatAdam emits torch.optim.Adam's state layout)."""
    ckpt = {"epoch": int(epoch), "global_step": int(global_step), "pytorch-lightning_version": CKPT_VERSION,
            "state_dict": {k: v.detach().cpu().clone() for k, v in module.state_dict().items()},
            "optimizer_states": [o.state_dict() for o in _optimizers_of(module)],
            "lr_schedulers": [], "hyper_parameters": dict(getattr(module, "hparams", {}) or {})}
    if extra:
        ckpt.update(extra)
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    torch.save(ckpt, path)
    return path


def restore_checkpoint(module, path, strict=True):
    """Loads weights, optimiser moments / step counts and the epoch counter from a Lightning-format checkpoint written
    by this package OR by the reference under Lightning (torch.optim.Adam states).  Returns the checkpoint dict."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    module.load_state_dict(ckpt["state_dict"], strict=strict)
    opts = _optimizers_of(module)
    states = ckpt.get("optimizer_states", [])
    assert len(states) in (0, len(opts)), f"checkpoint holds {len(states)} optimizer states, module has {len(opts)}"
    for o, s in zip(opts, states):
        o.load_state_dict(s)
        if hasattr(o, "invalidate_shadow"):
            o.invalidate_shadow()
    module.current_epoch = int(ckpt.get("epoch", 0))
    return ckpt


class Trainer:
    """fit(module, batches) with the reference Trainer's knobs that matter on the hot path.
    batches: an iterable of batch dicts (host or device tensors) or a callable step -> batch; the loop runs one
    optimisation step per batch: `loss = training_step(batch, i); backward; optimizer.step()` for automatic optimisation
    (PlayLMP), `training_step(batch)` alone for manual optimisation (TACORL steps its own optimisers).
    Under torchrun (WORLD_SIZE > 1) every FlatAdam averages its gradient over the ranks (NCCL), overlapped with the
    encoder backward; rank r must be fed its own shard (parallel.shard_batch)."""

    def __init__(self, max_steps=-1, max_epochs=1, precision="bf16", use_cuda_graph=True, default_root_dir=None,
                 steps_per_epoch=None, save_every_n_epochs=1, device=None):
        self.max_steps, self.max_epochs = max_steps, max_epochs
        self.precision = {"16": "bf16", 16: "bf16", "bf16": "bf16", "32": "fp32", 32: "fp32", "fp32": "fp32"}[precision]
        self.use_cuda_graph = use_cuda_graph
        self.root = Path(default_root_dir) if default_root_dir else None
        self.steps_per_epoch = steps_per_epoch
        self.save_every_n_epochs = save_every_n_epochs
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device or torch.device("cuda", self.local)
        self.global_step = 0
        self.losses = []

    def _setup(self, module):
        torch.cuda.set_device(self.device)
        ops.set_precision(self.precision)
        if self.world > 1 and not dist.is_initialized():
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(parallel.collective_channels(self.world)))
            os.environ.setdefault("NCCL_MIN_NCHANNELS", str(parallel.collective_channels(self.world)))
            dist.init_process_group("nccl", device_id=self.device)
        module.to(self.device)
        module.train()
        opts = _optimizers_of(module)
        if self.world > 1:
            for o in opts:
                parallel.attach_data_parallel(o, self.world)
        if getattr(module, "automatic_optimization", True):
            return runtime.play_lmp_step_fn(module, opts[0])
        return runtime.tacorl_step_fn(module)

    def fit(self, module, batches, ckpt_path=None):
        step_fn = self._setup(module)
        if ckpt_path is not None and Path(ckpt_path).is_file():
            ckpt = restore_checkpoint(module, ckpt_path)
            self.global_step = int(ckpt.get("global_step", 0))
        get = batches if callable(batches) else None
        it = None if get else iter(batches)
        graphed = None
        epoch0 = int(getattr(module, "current_epoch", 0))
        done_in_epoch = 0
        while True:
            if self.max_steps >= 0 and self.global_step >= self.max_steps:
                break
            if module.current_epoch - epoch0 >= self.max_epochs and self.max_steps < 0:
                break
            try:
                batch = get(self.global_step) if get else next(it)
            except StopIteration:
                break
            if batch is None:
                break
            if self.use_cuda_graph:
                if graphed is None:
                    graphed = runtime.GraphedTrainStep(step_fn, batch, device=self.device, warmup=2)
                out = graphed(batch)
            else:
                out = step_fn(_to_device(batch, self.device))
            if torch.is_tensor(out):
                self.losses.append(out.detach().clone())
            self.global_step += 1
            done_in_epoch += 1
            if self.steps_per_epoch and done_in_epoch >= self.steps_per_epoch:
                done_in_epoch = 0
                module.current_epoch += 1          # (a graph captured for the BC epochs is swapped when bc_epochs is crossed)
                if self.root and self.rank == 0 and module.current_epoch % self.save_every_n_epochs == 0:
                    save_checkpoint(module, self.root / "saved_models" / f"tacorl_epoch_{module.current_epoch:02d}_.ckpt",
                                    module.current_epoch, self.global_step)
        if self.root and self.rank == 0:
            save_checkpoint(module, self.root / "saved_models" / "last.ckpt", module.current_epoch, self.global_step)
        torch.cuda.synchronize(self.device)
        return self


try:  # pragma: no cover - pytorch_lightning is not in this image
    import pytorch_lightning as pl

    class B200Strategy(pl.strategies.SingleDeviceStrategy):
        """`pl.Trainer(strategy=B200Strategy(), devices=1, ...)` per torchrun rank: Lightning keeps its loops, logging and
        checkpoint callbacks; the gradient exchange is tacorl_b200.parallel's NCCL all-reduce inside FlatAdam.step()."""

        def __init__(self):
            local = int(os.environ.get("LOCAL_RANK", "0"))
            super().__init__(device=torch.device("cuda", local))
            self._world = int(os.environ.get("WORLD_SIZE", "1"))

        def setup(self, trainer):
            if self._world > 1 and not dist.is_initialized():
                dist.init_process_group("nccl", device_id=self.root_device)
            super().setup(trainer)
            if self._world > 1:
                for o in trainer.optimizers:
                    parallel.attach_data_parallel(o, self._world)
except Exception:
    B200Strategy = None
