"""N > 1 host logic on the CPU: world_size 2 and 3 over gloo (slab partition + ghost import plan)."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world,dim,p,reps", [(2, 3, 3, [7, 8, 20]), (3, 2, 3, [9, 20]), (2, 1, 5, [40]), (2, 3, 5, [6, 6, 23])])
def test_ghost_import_plan_over_gloo(lib, world, dim, p, reps):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "mp_halo_worker.py"), str(r), str(world), str(port),
                               str(dim), str(p)] + [str(n) for n in reps], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(world)]
    outs = [p_.communicate(timeout=300)[0] for p_ in procs]
    for p_, o in zip(procs, outs):
        assert p_.returncode == 0, o
        assert "OK" in o


@pytest.mark.parametrize("world,dim,p,n1", [(2, 2, 3, 20), (3, 2, 3, 24), (2, 3, 3, 10)])
def test_cut_rows_per_rank_over_gloo(lib, world, dim, p, n1):
    """BASELINE configuration 5's data path on the CPU: every rank generates the cut rows of its slab with no
    communication; after the library's ghost import (add_ghost_layer = 1) tensor-product rows + attached rows applied
    to locally stored data reproduce the one-rank oracle matrix on the owned range."""
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "mp_cut_worker.py"), str(r), str(world), str(port),
                               str(dim), str(p), str(n1)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = [p_.communicate(timeout=600)[0] for p_ in procs]
    for p_, o in zip(procs, outs):
        assert p_.returncode == 0, o
        assert "OK" in o
