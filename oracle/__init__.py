"""CPU oracle for the HELEN call_consensus/predict hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``helen_b200/`` may import this package.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs (as the checker / the timed CPU arm,
never as the product).

Parity status: the reference (kishwarshafin/helen @ a075e9f) ships no tests, golden
vectors or trained checkpoints for this path, so parity is *unpinned by the
reference's own tests*.  It is pinned instead by fixtures minted from the reference's
own ``TransducerGRU`` class imported from ``/root/reference`` in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``); ``tests/test_oracle.py``
checks both oracle implementations against those fixtures.
"""
from .explicit import (  # noqa: F401
    OracleWeights,
    forward_chunk,
    predict_windows,
    chunk_starts,
    STATE_DICT_KEYS,
    state_dict_shapes,
)
from .torch_port import TransducerPort, predict_port, random_state_dict  # noqa: F401
