"""CPU oracle for the Rainbow-IQN Ape-X learner hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never as the thing measured or shipped.  The product
package (``rainbow_iqn_apex_b200``) never imports this package and fails loudly
when its CUDA library is missing.

What it is: a from-scratch restatement (torch fp32 on CPU for the network and
losses, numpy float64 for the sum-tree and replay assembly) of the reference
algorithm, each function citing the ``/root/reference`` file:line it follows.

Parity pin: the reference ships no tests and no golden vectors,
so the oracle is pinned against *outputs of the reference itself*: the script
``oracle/make_golden.py`` imports the unmodified reference modules from
``/root/reference`` (possible only in the dev container), runs them with injected
noise / quantiles / sample values, asserts this restatement reproduces them, and
writes the small fixtures under ``tests/golden/`` that travel to the GPU box.
"""
