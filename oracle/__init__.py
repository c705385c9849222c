"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference hot path (OSQP-style ADMM + the condensed-QP
builders of /root/reference/src/ModelPredictiveControlAPI.cpp).  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package; the product (``solvempc_b200``) never does.
"""
from .oracle import *  # noqa: F401,F403
