"""CPU oracle for the InvertAvatar generator-forward hot path.

TEST INFRASTRUCTURE ONLY.  This package is a from-scratch CPU (torch fp32) restatement of
the reference algorithm; every function cites the reference file:line it follows.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU baseline.  The
product package ``invertavatar_b200`` never imports it and has no CPU fallback.

Pinning: the reference ships no tests, golden vectors or fixtures, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, imported unmodified from
``/root/reference`` on CPU in the build container by ``tests/golden/make_golden.py``; the
resulting vectors are committed under ``tests/golden/`` and ``tests/test_oracle_golden.py``
checks the oracle against them on every run (no GPU needed).

Third-party arithmetic: conv2d / conv_transpose2d / grid_sample / antialiased interpolate /
sort / searchsorted come from PyTorch ATen in both the reference (pinned pytorch=1.11.0,
``environment.yml:60``) and here (torch 2.11); ``cv2.floodFill`` (opencv-python 4.6.0.66,
``environment.yml:131``) is restated in numpy in ``triplane.fill_mouth``.
"""
