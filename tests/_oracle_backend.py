"""TEST-ONLY: run the engine's op lists through the NumPy oracle on host buffers, so that the
fuser / partitioner / exchange logic can be exercised without a GPU.  The product never does
this (ramba_b200.runtime raises without CUDA)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def install():
    from oracle import vm
    from ramba_b200.runtime import RT

    RT.set_test_executor(vm.run_deferred_ops, vm.reduce_partials, device="cpu")
