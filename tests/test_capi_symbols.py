"""The C-ABI library loads and exports every symbol include/avg_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from assistive_vr_gym_b200.build import build
    return ctypes.CDLL(build())


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "avg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(avg_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/avg_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from assistive_vr_gym_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_layout_constants_match(lib):
    from assistive_vr_gym_b200.compiler import blob
    assert lib.avg_env_stride() == blob.ENV_STRIDE == 192


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.avg_create(0, 4, ctypes.byref(h))
    assert rc != 0 and not h.value
    lib.avg_last_error.restype = ctypes.c_char_p
    assert b"no CUDA device" in lib.avg_last_error(None)


def test_make_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from assistive_vr_gym_b200 import make, capi
    with pytest.raises(capi.AvgError):
        make("ScratchItchJaco-v0", num_envs=2)


def test_struct_sizes_agree_between_c_and_numpy(oracles):
    from assistive_vr_gym_b200.compiler import blob
    from assistive_vr_gym_b200.capi import CONTACT_DT
    sizes = oracles[0].sizes()
    from assistive_vr_gym_b200.compiler.reset import RESET_TABLE_DT
    assert sizes == [blob.HEADER_DT.itemsize, blob.BODY_DT.itemsize, blob.DOF_DT.itemsize, blob.SHAPE_DT.itemsize,
                     blob.FRAME_DT.itemsize, CONTACT_DT.itemsize, RESET_TABLE_DT.itemsize]
