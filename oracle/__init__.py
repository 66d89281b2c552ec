"""CPU oracle for the UnSAMFlow PWC-Lite hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; ``unsamflow_b200`` never does.

Two checkers live here:

* ``oracle.c_oracle``  -- numpy front end of ``usf_oracle.c`` (plain C restatement,
  double accumulation; pinned against the reference through ``tests/golden``).
* ``oracle.torch_port`` -- the same path restated through the PyTorch ops the
  reference itself calls (``F.pad``/``mean``, ``F.grid_sample``); it is what the
  reference executes on a CPU, so it is the timed ``cpu_baseline`` ("port").

Parity pinning status: the reference ships no golden vectors for this path; both
checkers are pinned against outputs of the reference generated in the build
container by ``oracle/make_golden.py`` (fixtures in ``tests/golden``).
"""
from . import c_oracle  # noqa: F401
