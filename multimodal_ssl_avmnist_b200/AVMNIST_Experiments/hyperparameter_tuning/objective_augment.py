"""`process_augment_config` with the reference's signature (hyperparameter_tuning/objective_augment.py:8-96).  The Optuna
search driver itself is outside the hot-path scope; the trial branch only needs an object with suggest_float / suggest_int /
suggest_categorical, so a real optuna.Trial still works."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _compat  # noqa: F401,E402
from multimodal_ssl_avmnist_b200.augment import values_from_config  # noqa: E402


def process_augment_config(trial, config, is_hyperparameter_search=True):
    """-> {'augmentations': {view: {aug: args}}, 'augmentation_probabilities': {view: {aug: p}}}"""
    if not is_hyperparameter_search:
        return values_from_config(config)
    out = {"augmentations": {"global_views": {}, "local_views": {}}, "augmentation_probabilities": {"global_views": {}, "local_views": {}}}
    for view in ("global_views", "local_views"):
        for aug, space in config["optuna"]["augmentations"][view].items():
            args = {}
            for name, info in space.items():
                key = f"{view}.{aug}.{name}"
                if name == "p":
                    out["augmentation_probabilities"][view][aug] = trial.suggest_float(key, info["low"], info["high"])
                elif info["type"] == "uniform":
                    args[name] = trial.suggest_float(key, info["low"], info["high"])
                elif info["type"] == "int":
                    args[name] = trial.suggest_int(key, info["low"], info["high"], step=info.get("step", 1))
                elif info["type"] == "categorical":
                    args[name] = trial.suggest_categorical(key, info["choices"])
                else:
                    raise ValueError(f"Unknown parameter type: {info['type']} for {name}")
            if args:
                out["augmentations"][view][aug] = args
    return out


def objective(trial, config, model_dir, model):
    raise NotImplementedError("the Optuna augmentation search is outside the B200 hot-path scope (DESIGN.md section 1)")
