"""Module alias so that the reference's `from loss import OGMFlow_loss` (train.py:6) maps onto this package."""
from .evaluation import OGMFlow_loss  # noqa: F401
