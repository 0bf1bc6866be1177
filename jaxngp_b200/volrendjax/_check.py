"""Shape/dtype contract helpers (the role chex plays in the reference's abstract.py files)."""
import torch


def assert_shape(t, shape, name):
    if tuple(t.shape) != tuple(shape):
        raise AssertionError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")


def require_f32(t, op, what):
    if t.dtype != torch.float32:
        raise NotImplementedError(f"{op} is only implemented for {what} of `float32` type, got {t.dtype}")


def as_u32(t, name):
    """uint32 arrays travel as int32 (same bits)."""
    if t.dtype == torch.int32:
        return t
    if t.dtype == torch.uint32:
        return t.view(torch.int32)
    raise NotImplementedError(f"{name}: expected a uint32 array (carried as torch.int32/uint32), got {t.dtype}")


def positive(x, name):
    if not x > 0:
        raise AssertionError(f"{name} must be positive, got {x}")


def non_negative(x, name):
    if not x >= 0:
        raise AssertionError(f"{name} must be non-negative, got {x}")
