"""A small stand-in for the part of `lightning.pytorch` the reference's training path uses (run_dino.py:191-225,
351-389; models/dino.py LightningModules).  The API mirror imports the real package when it is installed and this
module otherwise, so `run_dino.py` runs without a Lightning install.  Single-process-per-GPU: with `strategy="ddp"` the
Trainer expects to be launched under torchrun (one process per GPU) and lets the step engine do the NCCL exchange.

Implemented surface: LightningModule (log, save_hyperparameters, hparams, device, load_from_checkpoint, hooks),
LightningDataModule, Callback, Trainer(max_epochs, max_steps, devices, strategy, precision, log_every_n_steps, logger,
callbacks, deterministic, limit_train_batches).fit / save_checkpoint / callback_metrics, seed_everything,
callbacks.ModelCheckpoint / EarlyStopping, loggers.CSVLogger.
"""
import csv
import inspect
import os
import random
import types

import numpy as np
import torch


def seed_everything(seed, workers=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    os.environ["PL_GLOBAL_SEED"] = str(seed)
    return seed


class LightningModule(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self._trainer = None
        self._hparams = {}

    # -- hyper-parameters / checkpoints --------------------------------------------------------------------
    def save_hyperparameters(self, *args, ignore=None, **kwargs):
        frame = inspect.currentframe().f_back
        init_args = {}
        while frame is not None:                      # walk up through subclass __init__ frames
            loc = frame.f_locals
            if loc.get("self") is self and frame.f_code.co_name == "__init__":
                for k, v in loc.items():
                    if k not in ("self", "__class__", "args", "kwargs") and not k.startswith("_"):
                        init_args.setdefault(k, v)
                for k, v in (loc.get("kwargs") or {}).items():
                    init_args.setdefault(k, v)
            frame = frame.f_back
        ignore = set(ignore or [])
        self._hparams = {k: v for k, v in init_args.items() if k not in ignore}

    @property
    def hparams(self):
        return types.SimpleNamespace(**self._hparams)

    @classmethod
    def load_from_checkpoint(cls, path, map_location=None, **overrides):
        ckpt = torch.load(path, map_location=map_location or "cpu", weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(overrides)
        sig = inspect.signature(cls.__init__).parameters
        accepts_kwargs = any(p.kind == inspect.Parameter.VAR_KEYWORD for p in sig.values())
        model = cls(**{k: v for k, v in hp.items() if accepts_kwargs or k in sig})
        model.load_state_dict(ckpt["state_dict"])
        return model

    # -- runtime ------------------------------------------------------------------------------------------
    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    @property
    def trainer(self):
        return self._trainer

    @property
    def current_epoch(self):
        return self._trainer.current_epoch if self._trainer else 0

    @property
    def global_step(self):
        return self._trainer.global_step if self._trainer else 0

    def log(self, name, value, on_step=None, on_epoch=None, prog_bar=False, **kw):
        if self._trainer is not None:
            self._trainer._log(name, value, on_step=on_step, on_epoch=on_epoch)

    def log_dict(self, d, **kw):
        for k, v in d.items():
            self.log(k, v, **kw)

    def configure_optimizers(self):
        raise NotImplementedError

    def training_step(self, batch, batch_idx):
        raise NotImplementedError

    def on_train_epoch_end(self):
        pass

    def on_train_start(self):
        pass


class LightningDataModule:
    def __init__(self):
        pass

    def prepare_data(self):
        pass

    def setup(self, stage=None):
        pass


class Callback:
    def on_train_start(self, trainer, pl_module): pass
    def on_train_end(self, trainer, pl_module): pass
    def on_train_epoch_start(self, trainer, pl_module): pass
    def on_train_epoch_end(self, trainer, pl_module): pass
    def on_train_batch_start(self, trainer, pl_module, batch, batch_idx): pass
    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx): pass


class ModelCheckpoint(Callback):
    def __init__(self, dirpath=None, monitor=None, save_top_k=1, mode="min", filename=None, **kw):
        self.dirpath, self.monitor, self.mode, self.save_top_k = dirpath, monitor, mode, save_top_k
        self.best_model_path, self.best_model_score = "", None

    def on_train_epoch_end(self, trainer, pl_module):
        if self.dirpath is None or self.save_top_k == 0 or trainer.global_rank != 0:
            return
        score = trainer.callback_metrics.get(self.monitor) if self.monitor else None
        score = float(score) if score is not None else None
        better = (self.best_model_score is None or score is None
                  or (score > self.best_model_score if self.mode == "max" else score < self.best_model_score))
        if better:
            os.makedirs(self.dirpath, exist_ok=True)
            path = os.path.join(self.dirpath, f"epoch={trainer.current_epoch}-step={trainer.global_step}.ckpt")
            trainer.save_checkpoint(path)
            if self.best_model_path and os.path.exists(self.best_model_path) and self.best_model_path != path:
                os.remove(self.best_model_path)
            self.best_model_path, self.best_model_score = path, score


class EarlyStopping(Callback):
    def __init__(self, monitor=None, patience=3, mode="min", **kw):
        self.monitor, self.patience, self.mode = monitor, patience, mode
        self.best, self.bad = None, 0

    def on_train_epoch_end(self, trainer, pl_module):
        v = trainer.callback_metrics.get(self.monitor)
        if v is None:
            return
        v = float(v)
        if self.best is None or (v > self.best if self.mode == "max" else v < self.best):
            self.best, self.bad = v, 0
        else:
            self.bad += 1
            if self.bad >= self.patience:
                trainer.should_stop = True


class CSVLogger:
    """Writes <save_dir>/<name>/version_<n>/metrics.csv with one row per logging event (columns as Lightning names them:
    epoch, step, <metric>_step, <metric>_epoch, ...)."""

    def __init__(self, save_dir, name="lightning_logs", version=None, **kw):
        self.save_dir, self.name = save_dir, name
        root = os.path.join(save_dir, name)
        if version is None:
            version = 0
            while os.path.exists(os.path.join(root, f"version_{version}")):
                version += 1
        self.version = version
        self.log_dir = os.path.join(root, f"version_{version}")
        self.rows = []

    def log_metrics(self, metrics, step, epoch):
        row = {"epoch": epoch, "step": step}
        row.update({k: (float(v) if v is not None else None) for k, v in metrics.items()})
        self.rows.append(row)

    def save(self):
        if not self.rows:
            return
        os.makedirs(self.log_dir, exist_ok=True)
        cols = ["epoch", "step"] + sorted({k for r in self.rows for k in r} - {"epoch", "step"})
        with open(os.path.join(self.log_dir, "metrics.csv"), "w", newline="") as f:
            wr = csv.DictWriter(f, fieldnames=cols)
            wr.writeheader()
            wr.writerows(self.rows)


callbacks = types.SimpleNamespace(ModelCheckpoint=ModelCheckpoint, EarlyStopping=EarlyStopping, Callback=Callback)
loggers = types.SimpleNamespace(CSVLogger=CSVLogger)


class Trainer:
    def __init__(self, max_epochs=1, max_steps=-1, devices="auto", strategy="auto", precision=None, log_every_n_steps=50, logger=None,
                 callbacks=None, deterministic=False, limit_train_batches=None, accelerator="auto", enable_progress_bar=False, **kw):
        self.max_epochs, self.max_steps = max_epochs, max_steps
        self.strategy, self.precision = strategy, precision
        self.log_every_n_steps = max(1, log_every_n_steps)
        self.logger = logger
        self.callbacks = list(callbacks or [])
        self.limit_train_batches = limit_train_batches
        self.callback_metrics = {}
        self.current_epoch, self.global_step = 0, 0
        self.should_stop = False
        self.global_rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self._epoch_acc = {}
        self.model = None
        self._dist_sampler = None

    # -- logging ------------------------------------------------------------------------------------------
    def _log(self, name, value, on_step=None, on_epoch=None):
        v = value.detach() if isinstance(value, torch.Tensor) else value
        in_step = getattr(self, "_in_step", False)
        if on_step is None:
            on_step = in_step
        if on_epoch is None:
            on_epoch = not in_step
        both = on_step and on_epoch
        if on_step:
            key = f"{name}_step" if both else name
            self.callback_metrics[key] = v
            self.callback_metrics[name] = v
            if self.global_step % self.log_every_n_steps == 0:
                self._pending_step[key] = v
        if on_epoch:
            if in_step:
                acc = self._epoch_acc.setdefault(name, [0.0, 0, both])
                acc[0] = acc[0] + v
                acc[1] += 1
            else:
                self.callback_metrics[name] = v
                self._pending_epoch[name] = v

    def _flush(self, pending):
        if pending and self.logger is not None and self.global_rank == 0:
            self.logger.log_metrics({k: (float(v) if v is not None else None) for k, v in pending.items()}, self.global_step, self.current_epoch)
        pending.clear()

    def _shard_loader(self, loader):
        """strategy='ddp' under torchrun: every rank must train on its own shard (real Lightning injects a DistributedSampler,
        run_dino.py:359).  Loaders that can shard themselves (DeviceResidentLoader.set_rank_shard) are told their rank; a plain
        torch DataLoader is rebuilt around a DistributedSampler; anything else is refused rather than silently replicated."""
        if self.world_size <= 1 or loader is None:
            return loader
        if hasattr(loader, "set_rank_shard"):
            loader.set_rank_shard(self.global_rank, self.world_size)
            return loader
        if isinstance(loader, torch.utils.data.DataLoader):
            if isinstance(loader.sampler, torch.utils.data.distributed.DistributedSampler):
                return loader
            shuffle = isinstance(loader.sampler, torch.utils.data.RandomSampler)
            sampler = torch.utils.data.distributed.DistributedSampler(loader.dataset, num_replicas=self.world_size, rank=self.global_rank,
                                                                      shuffle=shuffle, seed=int(os.environ.get("PL_GLOBAL_SEED", "0")))
            self._dist_sampler = sampler
            return torch.utils.data.DataLoader(loader.dataset, batch_size=loader.batch_size, sampler=sampler, num_workers=loader.num_workers,
                                               collate_fn=loader.collate_fn, pin_memory=loader.pin_memory, drop_last=loader.drop_last)
        raise RuntimeError("Trainer(strategy='ddp'): cannot shard this train loader over the ranks (no set_rank_shard, not a DataLoader)")

    # -- fit ----------------------------------------------------------------------------------------------
    def fit(self, model, datamodule=None, train_dataloaders=None):
        self.model = model
        model._trainer = self
        self.datamodule = datamodule
        self._pending_step, self._pending_epoch = {}, {}
        if torch.cuda.is_available():         # before the loaders are built: a device-resident loader lives on THIS rank's GPU
            dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
            torch.cuda.set_device(dev)
            if self.strategy == "ddp" and self.world_size > 1 and not torch.distributed.is_initialized():
                torch.distributed.init_process_group("nccl", device_id=dev)
            model.to(dev)
        if datamodule is not None:
            datamodule.prepare_data()
            datamodule.setup("fit")
            loader = datamodule.train_dataloader()
        else:
            loader = train_dataloaders
        loader = self._shard_loader(loader)
        cfg = model.configure_optimizers()
        sched = None
        if isinstance(cfg, dict):
            opt = cfg["optimizer"]
            sc = cfg.get("lr_scheduler")
            sched = sc["scheduler"] if isinstance(sc, dict) else sc
        elif isinstance(cfg, (tuple, list)):
            opt = cfg[0][0] if isinstance(cfg[0], (list, tuple)) else cfg[0]
            sched = (cfg[1][0] if isinstance(cfg[1], (list, tuple)) else cfg[1]) if len(cfg) > 1 and cfg[1] else None
        else:
            opt = cfg
        self.optimizers = [opt]
        use_scaler = str(self.precision) in ("16-mixed", "16") and getattr(model, "wants_grad_scaler", False)
        scaler = torch.amp.GradScaler("cuda", enabled=use_scaler) if use_scaler else None
        model.train()
        model.on_train_start()
        for cb in self.callbacks:
            cb.on_train_start(self, model)
        for epoch in range(self.max_epochs):
            self.current_epoch = epoch
            self._epoch_acc = {}
            for cb in self.callbacks:
                cb.on_train_epoch_start(self, model)
            if getattr(self, "_dist_sampler", None) is not None:
                self._dist_sampler.set_epoch(epoch)
            for batch_idx, batch in enumerate(loader):
                if self.limit_train_batches is not None and batch_idx >= self.limit_train_batches:
                    break
                for cb in self.callbacks:
                    cb.on_train_batch_start(self, model, batch, batch_idx)
                self._in_step = True
                loss = model.training_step(batch, batch_idx)
                self._in_step = False
                opt.zero_grad(set_to_none=True)
                if scaler is not None:
                    scaler.scale(loss).backward()
                    scaler.step(opt)
                    scaler.update()
                else:
                    loss.backward()
                    opt.step()
                self.global_step += 1
                self._flush(self._pending_step)
                for cb in self.callbacks:
                    cb.on_train_batch_end(self, model, loss, batch, batch_idx)
                if 0 < self.max_steps <= self.global_step:
                    self.should_stop = True
                    break
            for name, (tot, n, both) in self._epoch_acc.items():
                if n:
                    key = f"{name}_epoch" if both else name
                    self.callback_metrics[key] = tot / n
                    self._pending_epoch[key] = tot / n
            if sched is not None:
                sched.step()
            model.on_train_epoch_end()
            for cb in self.callbacks:
                cb.on_train_epoch_end(self, model)
            self._flush(self._pending_epoch)
            if self.should_stop:
                break
        for cb in self.callbacks:
            cb.on_train_end(self, model)
        if self.logger is not None and self.global_rank == 0:
            self.logger.save()
        return self

    def save_checkpoint(self, path):
        if self.global_rank != 0:
            return
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save({"state_dict": {k: v.detach().cpu() for k, v in self.model.state_dict().items()}, "epoch": self.current_epoch,
                    "global_step": self.global_step, "hyper_parameters": dict(getattr(self.model, "_hparams", {}))}, path)
