"""oracle/ -- TEST INFRASTRUCTURE.  CPU checkers for the PyCPET hot path.

Nothing in ``pycpet_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` leg / ``--impl reference``)
may import it, and only as the checker or the timed CPU baseline.

* ``oracle.f64``  -- float64 restatement (``cpet_oracle.c``), ground truth for fields, ESP,
  streamlines.
* ``oracle.hist`` -- NumPy restatement of ``np.histogram2d`` binning + the chi^2 distance.
* ``oracle.ref``  -- ctypes binding to ``oracle/_ref/math_module_v{3,4}.so``, i.e. the
  reference's own ``math_module.c`` compiled in place (``make -C oracle ref``).
"""
