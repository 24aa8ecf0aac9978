"""AVMNIST data path, B200-native: datasets / data modules with the reference's names and constructor arguments
(utils/get_data.py), and `MultiModalAugmentation`, whose multi-crop views are produced by the fused CUDA augmentation
kernels (libavmnist_b200.so) instead of per-sample torchvision / torchaudio calls in DataLoader workers.

Two ways to get views:
  * `MultiModalAugmentation.__call__(image [1,28,28], audio [1,112,112])` -- the reference's per-sample call, same return
    shapes; parameters are drawn on the host from torch's CPU generator / Python's `random` in the reference's order, so an
    identically seeded call reproduces the reference's crops, rotations, masks and noise; the pixels come from the CUDA
    kernels (B = 1 launch).  Needs a CUDA device; use num_workers=0.
  * the fast path: the data modules yield un-augmented batches and the Lightning module augments on the device with
    device-sampled parameters (`device_augmentation=True`, the default here).
"""
import os
import sys

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, random_split

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _compat import pl  # noqa: E402
from multimodal_ssl_avmnist_b200 import augment as A  # noqa: E402


# ---- the three custom transforms of the reference, as parameter holders -------------------------------------
class GaussianNoise(torch.nn.Module):
    def __init__(self, std=0.1):
        super().__init__()
        self.std = std

    def spec(self, p):
        return A.OpSpec(A.SPEC_NOISE, p, self.std)


class TimeWarpWithStretch(torch.nn.Module):
    def __init__(self, min_factor=0.8, max_factor=1.2, target_length=112):
        super().__init__()
        self.min_factor, self.max_factor, self.target_length = min_factor, max_factor, target_length

    def spec(self, p):
        return A.OpSpec(A.SPEC_TIME_WARP, p, self.min_factor, self.max_factor)


class GroupedMasking(torch.nn.Module):
    def __init__(self, mask_ratio=0.5, group_size=4):
        super().__init__()
        self.mask_ratio, self.group_size = mask_ratio, group_size

    def spec(self, p):
        return A.OpSpec(A.SPEC_GROUP_MASK, p, int(self.mask_ratio * (112 // self.group_size) ** 2), self.group_size)


_SPEC_NAMES = {A.SPEC_RRC: "RandomResizedCrop", A.SPEC_ROTATE: "RandomRotation", A.SPEC_AFFINE: "RandomAffine", A.SPEC_ERASE: "RandomErasing",
               A.SPEC_FREQ_MASK: "FrequencyMasking", A.SPEC_TIME_MASK: "TimeMasking", A.SPEC_NOISE: "GaussianNoise",
               A.SPEC_GROUP_MASK: "GroupedMasking", A.SPEC_TIME_WARP: "TimeWarpWithStretch"}


class MultiModalAugmentation:
    def __init__(self, n_global_views=2, n_local_views=4, global_spec_size=112, local_spec_size=112, augment_values=None):
        if global_spec_size != 112 or local_spec_size != 112:
            raise ValueError("the compiled augmentation kernels handle 112x112 spectrograms")
        self.n_global_views, self.n_local_views = n_global_views, n_local_views
        self.global_spec_size, self.local_spec_size = global_spec_size, local_spec_size
        self.augment_values = augment_values
        ig, il = A.image_chains()
        ag, al = A.default_audio_chains() if augment_values is None else A.audio_chains_from_values(augment_values)
        self.global_transforms = {"image": ig, "audio": ag}
        self.local_transforms = {"image": il, "audio": al}
        self._sampler = A.HostSampler()

    @torch.no_grad()
    def __call__(self, images, audios):
        """images [1,28,28], audios [1,112,112] (CPU or CUDA) -> (gi [Vg,1,28,28], ga [Vg,1,112,112], li, la) on the CUDA device."""
        from multimodal_ssl_avmnist_b200 import ops
        if not torch.cuda.is_available():
            raise RuntimeError("MultiModalAugmentation runs its pixels on the GPU: no CUDA device available")
        dev = images.device if images.is_cuda else torch.device("cuda", torch.cuda.current_device())
        Vg, Vl = self.n_global_views, self.n_local_views
        V = Vg + Vl
        img_ops = np.zeros((1, V, A.MAX_OPS, A.OP_WORDS), dtype=np.int32)
        aud_ops = np.zeros_like(img_ops)
        bits = np.zeros((1, V, A.GROUP_WORDS), dtype=np.uint32)
        noise = torch.zeros(1, V, 112, 112)
        for v in range(V):                      # reference order: per view, image chain then audio chain
            ci, ca = (self.global_transforms["image"], self.global_transforms["audio"]) if v < Vg else \
                     (self.local_transforms["image"], self.local_transforms["audio"])
            o, _, _ = self._sampler.sample_view(ci, 28, 28)
            A.pack_ops(o, img_ops[0, v])
            o, gb, nz = self._sampler.sample_view(ca, 112, 112)
            A.pack_ops(o, aud_ops[0, v])
            bits[0, v] = A.pack_group_bits(gb)
            if nz is not None:
                noise[0, v] = nz
        out_i = torch.empty(V, 1, 28, 28, device=dev)
        out_a = torch.empty(V, 1, 112, 112, device=dev)
        ops.aug_apply_image(images.reshape(1, 28, 28).float().contiguous().to(dev), torch.from_numpy(img_ops).to(dev), out_i)
        ops.aug_apply_audio(audios.reshape(1, 112, 112).float().contiguous().to(dev), torch.from_numpy(aud_ops).to(dev),
                            torch.from_numpy(bits.view(np.int32)).to(dev), out_a, noise=noise.to(dev))
        return out_i[:Vg], out_a[:Vg], out_i[Vg:], out_a[Vg:]

    def __str__(self):
        def fmt(chain):
            return [f"      {_SPEC_NAMES[o.kind]}(p={'always' if o.p == A.ALWAYS else o.p}, args={[round(a, 6) for a in o.a]})" for o in chain]
        lines = ["MultiModalAugmentation(", f"  n_global_views={self.n_global_views},", f"  n_local_views={self.n_local_views},",
                 f"  global_spec_size={self.global_spec_size},", f"  local_spec_size={self.local_spec_size},", "  global_transforms:",
                 "    image: ["] + fmt(self.global_transforms["image"]) + ["    ],", "    audio: ["] + fmt(self.global_transforms["audio"]) + \
                ["    ]", "  local_transforms:", "    image: ["] + fmt(self.local_transforms["image"]) + ["    ],", "    audio: ["] + \
                fmt(self.local_transforms["audio"]) + ["    ]", ")"]
        return "\n".join(lines)


class SimCLRMultiModalAugmentation:
    """Two augmented views per modality for the SimCLR-style models (reference utils/get_data.py:299-408).  Like the reference, the
    transforms act on the whole [B,C,H,W] batch at once: ONE parameter set (crop box, angle, elastic field, blur sigma, masks,
    warp rate) per call and view, Gaussian noise per element.  Parameters are drawn on the host in the reference's RNG order
    (HostSampler), the pixels are produced by the CUDA augmentation kernels (crop-resize, affine, ElasticTransform, GaussianBlur /
    crop-resize, time-warp, frequency / time masks, noise).  `augment_values` is accepted and, as in the reference (whose custom
    list is built but never installed, :366-383), does not change the transforms."""

    def __init__(self, image_size=28, spec_size=112, augment_values=None):
        if image_size != 28 or spec_size != 112:
            raise ValueError("the compiled augmentation kernels handle 28x28 images and 112x112 spectrograms")
        self.image_size, self.spec_size = image_size, spec_size
        self.augment_values = augment_values
        self.image_transform, self.spectrogram_transform = A.simclr_chains()
        self._sampler = A.HostSampler()

    @torch.no_grad()
    def __call__(self, images, audios):
        """images [B,1,28,28], audios [B,1,112,112] -> (aug_images1, aug_audios1, aug_images2, aug_audios2), CUDA tensors of those shapes."""
        from multimodal_ssl_avmnist_b200 import ops
        if not torch.cuda.is_available():
            raise RuntimeError("SimCLRMultiModalAugmentation runs its pixels on the GPU: no CUDA device available")
        dev = images.device if images.is_cuda else torch.device("cuda", torch.cuda.current_device())
        B, V = images.shape[0], 2
        img_ops = np.zeros((B, V, A.MAX_OPS, A.OP_WORDS), dtype=np.int32)
        aud_ops = np.zeros_like(img_ops)
        grids = np.zeros((B, V, 2, 28, 28), dtype=np.float32)
        noise = torch.zeros(B, V, 112, 112)
        for v in range(V):                      # reference order: both image views, then both spectrogram views (:398-404)
            o, _, _ = self._sampler.sample_view(self.image_transform, 28, 28, batch=B)
            A.pack_ops(o, img_ops[0, v])
            img_ops[:, v] = img_ops[0, v]
            if self._sampler.last_grid is not None:
                grids[:, v] = self._sampler.last_grid
        for v in range(V):
            o, _, nz = self._sampler.sample_view(self.spectrogram_transform, 112, 112, batch=B)
            A.pack_ops(o, aud_ops[0, v])
            aud_ops[:, v] = aud_ops[0, v]
            if nz is not None:
                noise[:, v] = nz
        out_i = torch.empty(V, B, 28, 28, device=dev)
        out_a = torch.empty(V, B, 112, 112, device=dev)
        bits = torch.zeros(B, V, A.GROUP_WORDS, dtype=torch.int32, device=dev)
        ops.aug_apply_image(images.reshape(B, 28, 28).float().contiguous().to(dev), torch.from_numpy(img_ops).to(dev), out_i,
                            elastic_grid=torch.from_numpy(grids).to(dev))
        ops.aug_apply_audio(audios.reshape(B, 112, 112).float().contiguous().to(dev), torch.from_numpy(aud_ops).to(dev), bits, out_a,
                            noise=noise.to(dev))
        return out_i[0].unsqueeze(1), out_a[0].unsqueeze(1), out_i[1].unsqueeze(1), out_a[1].unsqueeze(1)


# ---- datasets -------------------------------------------------------------------------------------------------
class BaseAVMNISTDataset(Dataset):
    """image/<split>_data.npy (np.load-able, N x 784), audio/<split>_data_augmented_<type>.npy (headerless uint8 memmap,
    N x 112 x 112), <split>_labels.npy -- the on-disk layout of the reference (utils/get_data.py:427-436)."""

    def __init__(self, image_path, audio_path, labels_path, flatten_audio=False, flatten_image=False, unsqueeze_channel=True,
                 normalize_image=True, normalize_audio=True, compute_stats=False):
        self.labels = np.load(labels_path).astype(int)
        self.image_data = np.load(image_path, mmap_mode="r")
        self.audio_data = np.memmap(audio_path, dtype="uint8", mode="r", shape=(len(self.labels), 112, 112))
        self.flatten_audio, self.flatten_image, self.unsqueeze_channel = flatten_audio, flatten_image, unsqueeze_channel
        self.normalize_image, self.normalize_audio = normalize_image, normalize_audio
        self.audio_mean, self.audio_std = 0.0, 1.0
        if compute_stats and normalize_audio:
            a = np.asarray(self.audio_data, dtype=np.float64) / 255.0
            self.audio_mean, self.audio_std = float(a.mean(axis=(1, 2)).mean()), float(a.std(axis=(1, 2)).mean())

    def __len__(self):
        return len(self.labels)

    def _process_image_audio(self, idx):
        image = np.array(self.image_data[idx], dtype=np.float64)
        audio = np.array(self.audio_data[idx], dtype=np.float64)
        image = image if self.flatten_image else image.reshape(28, 28)
        audio = audio.reshape(-1) if self.flatten_audio else audio
        if self.normalize_image:
            image = image / 255.0
        if self.normalize_audio:
            audio = (audio / 255.0 - self.audio_mean) / self.audio_std
        if self.unsqueeze_channel:
            image, audio = image[None], audio[None]
        return image, audio


class AVMNISTDataset(BaseAVMNISTDataset):
    def __getitem__(self, idx):
        image, audio = self._process_image_audio(idx)
        return image, audio, torch.tensor(self.labels[idx], dtype=torch.long)


class AVMNISTSSLDataset(BaseAVMNISTDataset):
    """transform=None yields the un-augmented (image, audio) pair (device augmentation); otherwise the reference's views."""

    def __init__(self, *args, transform=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.transform = transform

    def _tensors(self, idx):
        image, audio = self._process_image_audio(idx)
        return torch.tensor(image, dtype=torch.float32), torch.tensor(audio, dtype=torch.float32)

    def __getitem__(self, idx):
        image, audio = self._tensors(idx)
        return (image, audio) if self.transform is None else tuple(t.cpu() for t in self.transform(image, audio))


class AVMNISTSSLDatasetExtended(AVMNISTSSLDataset):
    def __getitem__(self, idx):
        image, audio = self._tensors(idx)
        label = torch.tensor(self.labels[idx], dtype=torch.long)
        if self.transform is None:
            return image, audio, label
        return image, audio, label, tuple(t.cpu() for t in self.transform(image, audio))


# ---- data modules ---------------------------------------------------------------------------------------------
class BaseAVMNISTDataModule(pl.LightningDataModule):
    def __init__(self, data_dir, batch_size=128, num_workers=6, type="burst_noise", train_shuffle=True, flatten_audio=False,
                 flatten_image=False, unsqueeze_channel=True, normalize_image=True, normalize_audio=True, train_size=55000, val_size=5000,
                 test_size=10000):
        super().__init__()
        self.data_dir, self.batch_size, self.num_workers, self.type, self.train_shuffle = data_dir, batch_size, num_workers, type, train_shuffle
        self._ds_kwargs = dict(flatten_audio=flatten_audio, flatten_image=flatten_image, unsqueeze_channel=unsqueeze_channel,
                               normalize_image=normalize_image, normalize_audio=normalize_audio)
        self.train_size, self.val_size, self.test_size = train_size, val_size, test_size
        for split in ("train", "test"):
            setattr(self, f"{split}_image_path", f"{data_dir}image/{split}_data.npy")
            setattr(self, f"{split}_audio_path", f"{data_dir}audio/{split}_data_augmented_{type}.npy")
            setattr(self, f"{split}_labels_path", f"{data_dir}{split}_labels.npy")

    def prepare_data(self):
        for p in (self.train_image_path, self.train_audio_path, self.train_labels_path, self.test_image_path, self.test_audio_path,
                  self.test_labels_path):
            if not os.path.exists(p):
                raise FileNotFoundError(f"Data file not found: {p}")

    def _get_dataset_kwargs(self):
        return dict(self._ds_kwargs)

    def _train_cls(self):
        return AVMNISTDataset, {}

    def setup(self, stage=None):
        if stage in ("fit", None):
            cls, extra = self._train_cls()
            full = cls(self.train_image_path, self.train_audio_path, self.train_labels_path, **extra, **self._get_dataset_kwargs())
            n = len(full)
            tr = min(self.train_size, n - min(self.val_size, n // 10))
            self.train_dataset, self.val_dataset, _ = random_split(full, [tr, min(self.val_size, n - tr), n - tr - min(self.val_size, n - tr)])
        if stage in ("test", None):
            test = AVMNISTDataset(self.test_image_path, self.test_audio_path, self.test_labels_path, **self._get_dataset_kwargs())
            k = min(self.test_size, len(test))
            self.test_dataset, _ = random_split(test, [k, len(test) - k])

    def _loader(self, ds, shuffle):
        return DataLoader(ds, batch_size=self.batch_size, shuffle=shuffle, num_workers=self.num_workers,
                          persistent_workers=self.num_workers > 0, drop_last=shuffle)

    def train_dataloader(self):
        return self._loader(self.train_dataset, self.train_shuffle)

    def val_dataloader(self):
        return self._loader(self.val_dataset, False)

    def test_dataloader(self):
        return self._loader(self.test_dataset, False)


class AVMNISTDataModule(BaseAVMNISTDataModule):
    pass


class DeviceResidentLoader:
    """The whole training split resident in HBM: image fp32 [N,1,28,28] as stored on disk / 255, audio kept as the
    on-disk uint8 [N,1,112,112] (0.75 GB for 60k samples; the /255 of utils/get_data.py:467 happens inside the augmentation
    kernel), labels int64.  An epoch is a device-side permutation; a batch is one gather.  Yields (image, audio[, label]) device
    tensors -- the raw-batch form the training_step augments on the GPU -- so no DataLoader worker, pinned buffer or H2D copy
    sits on the step's critical path."""

    def __init__(self, dataset, indices, batch_size, device, shuffle=True, drop_last=True, with_labels=False, seed=0):
        base = dataset
        idx = np.sort(np.asarray(indices)) if indices is not None else np.arange(len(base))
        img = np.asarray(base.image_data[idx], dtype=np.float32).reshape(-1, 1, 28, 28)
        self.image = torch.from_numpy(img / 255.0 if base.normalize_image else img).to(device)
        self.audio = torch.from_numpy(np.ascontiguousarray(base.audio_data[idx])).reshape(-1, 1, 112, 112).to(device)      # uint8
        self.labels = torch.from_numpy(base.labels[idx].astype(np.int64)).to(device)
        self.batch_size, self.shuffle, self.drop_last, self.with_labels = batch_size, shuffle, drop_last, with_labels
        self.gen = torch.Generator(device=device).manual_seed(seed)
        self.device = device
        self.rank, self.world = 0, 1

    def set_rank_shard(self, rank, world):
        """Data parallel (what Lightning's DistributedSampler does for the reference under strategy='ddp', run_dino.py:359): every
        rank draws the SAME epoch permutation (same generator seed) and takes the strided slice rank::world of it, truncated so
        that all ranks see the same number of samples."""
        self.rank, self.world = int(rank), int(world)

    def _n_local(self):
        return self.image.shape[0] // self.world

    def __len__(self):
        n = self._n_local()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.image.shape[0]
        perm = torch.randperm(n, device=self.device, generator=self.gen) if self.shuffle else torch.arange(n, device=self.device)
        if self.world > 1:
            perm = perm[:self._n_local() * self.world][self.rank::self.world]
        for b in range(len(self)):
            sel = perm[b * self.batch_size:(b + 1) * self.batch_size]
            batch = (self.image.index_select(0, sel), self.audio.index_select(0, sel))
            yield batch + ((self.labels.index_select(0, sel),) if self.with_labels else ())


class AVMNISTDinoDataModule(BaseAVMNISTDataModule):
    """device_augmentation=True (default): batches are the un-augmented (image, audio) pairs and the multi-crop views are
    made on the GPU inside the training step; False: the reference's collated 4-tuple of views (needs num_workers=0)."""
    EXTENDED = False

    def __init__(self, data_dir, batch_size=32, num_workers=4, n_global_views=2, n_local_views=4, type="burst_noise", augmentations=None,
                 device_augmentation=True, device_resident=False):
        super().__init__(data_dir=data_dir, batch_size=batch_size, num_workers=num_workers, type=type)
        self.n_global_views, self.n_local_views = n_global_views, n_local_views
        self.augmentations = augmentations if augmentations is not None else MultiModalAugmentation(n_global_views, n_local_views)
        self.device_augmentation = device_augmentation
        self.device_resident = device_resident          # keep the training split in HBM (DeviceResidentLoader)
        if not device_augmentation and num_workers != 0:
            raise ValueError("per-sample CUDA augmentation in the dataset needs num_workers=0")

    def _train_cls(self):
        cls = AVMNISTSSLDatasetExtended if self.EXTENDED else AVMNISTSSLDataset
        return cls, {"transform": None if self.device_augmentation else self.augmentations}

    def get_view_config(self):
        return {"n_global_views": self.n_global_views, "n_local_views": self.n_local_views}

    def probe_dataloaders(self):
        """Labelled (image, audio, label) loaders over the train / validation subsets for the per-epoch linear probe."""
        def labelled(sub):
            ds = AVMNISTDataset.__new__(AVMNISTDataset)
            ds.__dict__.update(sub.dataset.__dict__)
            return torch.utils.data.Subset(ds, sub.indices)

        def collate(items):
            img = torch.tensor(np.stack([i[0] for i in items]), dtype=torch.float32)
            aud = torch.tensor(np.stack([i[1] for i in items]), dtype=torch.float32)
            return img, aud, torch.stack([i[2] for i in items])
        mk = lambda sub, sh: DataLoader(labelled(sub), batch_size=self.batch_size, shuffle=sh, num_workers=0, collate_fn=collate)
        return mk(self.train_dataset, True), mk(self.val_dataset, False)

    def train_dataloader(self):
        if self.device_resident and self.device_augmentation and torch.cuda.is_available():
            sub = self.train_dataset
            return DeviceResidentLoader(sub.dataset, sub.indices, self.batch_size, torch.device("cuda", torch.cuda.current_device()),
                                        shuffle=self.train_shuffle, with_labels=self.EXTENDED,
                                        seed=int(torch.initial_seed()) & 0x7FFFFFFF)       # the run's global seed: same permutation on every rank
        return super().train_dataloader()


class AVMNISTDinoDataModuleExtended(AVMNISTDinoDataModule):
    EXTENDED = True


def write_synthetic_avmnist(data_dir, n_train=512, n_test=128, type="burst_noise", seed=0):
    """AVMNIST-shaped synthetic files in the reference's on-disk layout (for smoke runs without the dataset)."""
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(data_dir, "image"), exist_ok=True)
    os.makedirs(os.path.join(data_dir, "audio"), exist_ok=True)
    for split, n in (("train", n_train), ("test", n_test)):
        np.save(os.path.join(data_dir, "image", f"{split}_data.npy"), rng.integers(0, 256, (n, 784)).astype(np.float64))
        mm = np.memmap(os.path.join(data_dir, "audio", f"{split}_data_augmented_{type}.npy"), dtype="uint8", mode="w+", shape=(n, 112, 112))
        mm[:] = rng.integers(0, 256, (n, 112, 112), dtype=np.uint8)
        mm.flush()
        np.save(os.path.join(data_dir, f"{split}_labels.npy"), rng.integers(0, 10, n))
