"""SLURM environment -> config['hardware'] (same keys as the reference's configs/update_config.py:3-14)."""
import os


def update_hardware_config(config):
    cpus = int(os.getenv("SLURM_CPUS_PER_TASK", 4))
    gpus = int(os.getenv("SLURM_GPUS", os.getenv("WORLD_SIZE", 0)))
    config.setdefault("hardware", {})
    config["hardware"].update(num_workers=cpus, device="gpu" if gpus > 0 else "cpu", num_gpus=gpus)
    print(f"Updated config: {config['hardware']}")
    return config
