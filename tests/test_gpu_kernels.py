"""Parity tests proper (-m gpu): every kernel through the C ABI against a PyTorch fp32 statement of
the same op / the CPU oracle, at the shapes of the BASELINE configs and at ragged edge cases.
The check bodies live in scripts/check_*.py so they can also be run standalone on a B200."""
import importlib.util
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _load(name):
    spec = importlib.util.spec_from_file_location(name, ROOT / "scripts" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("group", ["basic", "epilogue", "dgrad", "wgrad", "patch", "pair", "maps"])
def test_gemm(group):
    assert getattr(_load("check_gemm"), f"group_{group}")()


@pytest.mark.parametrize("group", ["rowops", "attn64", "attn32"])
def test_attention_fwd_and_row_kernels(group):
    assert getattr(_load("check_attn_rows"), f"group_{group}")()


@pytest.mark.parametrize("group", ["attnbwd", "adapters", "loss"])
def test_training_kernels(group):
    assert getattr(_load("check_train"), f"group_{group}")()
