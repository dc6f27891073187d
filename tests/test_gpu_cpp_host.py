"""Runs the C++ host-mirror self-test (tests/cpp/test_grid_search.cpp) on the GPU box: the PCL-shaped
pcc::search::GridSearch<PointT> class and the consumer drivers, through the C ABI, against an in-test brute force."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_grid_search")


@pytest.mark.gpu
def test_cpp_host_mirror():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout + out.stderr


def test_cpp_host_mirror_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    if not os.path.exists(EXE):
        pytest.skip("not built")
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0           # pcc::Error("... no CPU fallback") escapes main


ADAPTER = os.path.join(ROOT, "tests", "cpp", "test_pcl_adapter")


@pytest.mark.gpu
def test_pcl_adapter_branch_behind_search_base_pointer():
    """include/pcc/grid_search.hpp's PCC_HAVE_PCL branch (GridSearch derived from pcl::search::Search<PointT>), compiled against
    tests/cpp/pcl_stub and driven through a pcl::search::Search<PointT>::Ptr like a PCL consumer after setSearchMethod()."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    assert os.path.exists(ADAPTER), "run __graft_entry__.build() first"
    out = subprocess.run([ADAPTER], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "PCL adapter ok" in out.stdout, out.stdout + out.stderr


def test_pcl_adapter_branch_is_built():
    """The adapter branch compiles in this image (no PCL): __graft_entry__.build() makes tests/cpp/test_pcl_adapter."""
    if not os.path.exists(os.path.join(ROOT, "tests", "cpp", "test_grid_search")):
        pytest.skip("not built")
    assert os.path.exists(ADAPTER)
