"""CPU: the C-ABI library loads and exports every symbol include/avmnist_b200.h declares (no compute calls)."""
import ctypes
import os

import pytest

from multimodal_ssl_avmnist_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from multimodal_ssl_avmnist_b200.build import build
        build()
    return _lib.load()


def test_header_declares_functions():
    names = [n for n, _, _ in _lib.declared_functions()]
    assert len(names) >= 40 and len(set(names)) == len(names)
    for must in ("b200_ema_flat", "b200_dino_loss_fwd_bwd", "b200_aug_apply_audio", "b200_conv_fwd", "b200_linear_fwd",
                 "b200_infonce_fwd_bwd", "b200_adam_flat"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name, _, _ in _lib.declared_functions():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert lib.b200_abi_version() == _lib.ABI_VERSION


def test_argument_errors_are_reported_without_a_gpu(lib):
    # argument validation happens before any CUDA call, so it is checkable on a CPU-only box
    rc = lib.b200_ema_flat(None, None, 0, 0.996, 0.004, None)
    assert rc == -1 and b"ema_flat" in lib.b200_last_error()
    rc = lib.b200_conv_fwd(None, None, None, None, None, 0, 0, 1, 1, 1, 1, 1, 0, None)
    assert rc == -1
    assert lib.b200_conv_supported(8, 16, 56, 56, 5, 2) == 1
    assert lib.b200_conv_supported(3, 16, 56, 56, 5, 2) == 0
    assert lib.b200_dino_loss_parts(1) >= 1
    assert lib.b200_infonce_work_floats(8, 128) == 4 * 8 * 128 + 32
    assert lib.b200_conv_bwd_weight_work_floats(4, 8, 16, 56, 56, 5, 2) > 0
    assert lib.b200_conv_bwd_weight_work_floats(4, 3, 16, 56, 56, 5, 2) == -2


def test_ops_refuse_cpu_tensors(lib):
    import torch
    from multimodal_ssl_avmnist_b200 import ops
    with pytest.raises(_lib.B200Error):
        ops.ema_flat(torch.zeros(8), torch.zeros(8), 0.996)
