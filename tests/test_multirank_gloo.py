"""N>1 path on CPU: world sizes 2, 3, 4 and 8 over gloo.  Every rank runs the same driver program (SPMD),
owns one division of every array, exchanges operand pieces that cross ranks (halo planes, broadcast
operands, reduction partials) and must reproduce NumPy exactly."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, names):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ)
        env.update({"RANK": str(r), "WORLD_SIZE": str(world), "LOCAL_RANK": str(r), "MASTER_ADDR": "127.0.0.1",
                    "MASTER_PORT": str(port), "OMP_NUM_THREADS": "1"})
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_mr_worker.py"), names], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append((p.returncode, o))
    for rc, o in outs:
        assert rc == 0, o[-3000:]
    return outs


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_programs_multirank(world):
    outs = _run(world, "all")
    # the stencil / broadcast / axis-sum programs must really have crossed ranks
    assert any("bytes_sent=0" not in o for _, o in outs)
    # ... and the stage-2 reductions / operands needed everywhere went through collectives (all-reduce, all-gather)
    assert all("collectives=0 " not in o for _, o in outs)


@pytest.mark.timeout(300)
def test_local_border_halo_goes_into_the_ring():
    """Arrays created with local_border: the neighbours' edges are received into the ring of the padded block
    (getborder) and the stencil reads one buffer."""
    outs = _run(4, "sstencil_local_border")
    assert all("ring_receives=0 " not in o for _, o in outs), outs


@pytest.mark.timeout(300)
def test_distributed_file_load(tmp_path):
    """`load` of a file type that can be read in parts: every rank reads its own block (ramba/ramba.py:8930-8945,
    3929-3956); the values and a reduction over them match NumPy on every rank."""
    import numpy as onp

    path = str(tmp_path / "field.npy")
    onp.save(path, (onp.arange(46 * 61, dtype=onp.float64).reshape(46, 61) % 13) - 5.0)
    os.environ["RB200_MR_NPY"] = path
    try:
        _run(3, "load_npy")
    finally:
        del os.environ["RB200_MR_NPY"]
