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
