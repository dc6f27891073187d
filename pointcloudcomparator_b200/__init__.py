"""pointcloudcomparator_b200 -- B200-native batched kNN / radius search for PointCloudComparator's hot path.

Only what the path needs lives here: csrc/ (CUDA kernels + the C ABI of include/pcc/search.h, built into
libpcc_search.so), the ctypes binding, the Python host mirror of pcl::search::Search (search.GridSearch),
query sharding across GPUs (shard.py) and the seeded synthetic clouds of the BASELINE configs (synth.py).
"""
from . import synth  # noqa: F401

__all__ = ["GridSearch", "synth"]


def __getattr__(name):
    if name == "GridSearch":
        from .search import GridSearch
        return GridSearch
    raise AttributeError(name)
