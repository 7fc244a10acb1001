"""Host-side logic of the input prefetcher (slot accounting); the CUDA behaviour is exercised by bench.py's e2e leg
and tests/test_planner_gpu.py::test_prefetcher_matches_direct_copy."""
import pytest
import torch

from etpnav_b200 import pipeline


def test_prefetcher_needs_cuda_or_raises_cleanly():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(Exception):
        pipeline.HostInputPrefetcher("cuda:0")
