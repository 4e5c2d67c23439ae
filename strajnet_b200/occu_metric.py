"""Module alias so that the reference's `import occu_metric as occupancy_flow_metrics` (train.py:9) maps onto this package."""
from .evaluation import WaypointGrids, compute_occupancy_flow_metrics  # noqa: F401
