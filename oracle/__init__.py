"""
oracle -- CPU restatement of the reference's free-energy + gradient path.

TEST INFRASTRUCTURE ONLY.  Nothing under vgpa_b200/ imports this package; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs do, and there only as the checker (or as the CPU arm being timed).

Parity status: PINNED against outputs of the unmodified reference recorded in
tests/golden/*.npz (tests/test_oracle_golden.py).

    from oracle import Oracle, Problem
    prob = Problem(model="L96", method="rk2", D=40, N=1001, dt=0.01, ...)
    F, grad = Oracle().eval(prob, x)
"""
from .oracle import Oracle, Problem, prior_kl0, build_oracle, ORACLE_SO  # noqa: F401
