"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the ngp_pl Instant-NGP hot path (the `vren` extension and the
tiny-cuda-nn modules that ``ngp_pl/models/{rendering,custom_functions,networks}.py``
call).  PARITY UNPINNED: the reference ships neither the kernels' source
(``.gitignore:23-25``) nor any test or golden vector, so this
oracle is pinned only by (a) the shape/sentinel/in-place contracts at the
reference call sites, (b) analytic closed forms, and (c) agreement between its
two independent restatements: the serial C one in ``oracle/c/vren_oracle.c`` and
the vectorised torch one in ``oracle/vren_ref.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``google-nerf_b200``) never does.
"""
