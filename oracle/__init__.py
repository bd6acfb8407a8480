"""TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is the CPU checker for the CUDA hot path: a plain
PyTorch fp32 restatement of the reference algorithm (``hh_oracle.py``), the shim
that imports the real reference in the build container (``ref_import.py``) and
the script that pins the restatement against it (``make_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
``helping_hand_for_egocentric_videos_b200`` never does.
"""
