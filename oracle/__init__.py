"""CPU oracle for the AtmoNR hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU (numpy / torch-CPU fp32+fp64) restatement of the algorithms on the
north-star path of nasa/atmospheric-neural-rendering:

    ray sampling -> geodetic preprocess -> hash-grid / positional encoding -> density and
    radiance MLPs -> emission-absorption compositing -> per-band loss (+ backward, AdamW)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the *checker* (or as the timed CPU
baseline) -- never as the thing shipped.  The product path (``atmonr`` +
``libatmonr_b200.so``) does not import this package and fails loudly when its CUDA
library is missing.

Pinning status
--------------
* ``geodesy``, ``sampling``, ``rendering`` (incl. the losses) and ``nerf``: PINNED.  They are
  checked against golden vectors produced by importing the reference's own Python modules
  from ``/root/reference/src`` (``tests/golden/make_golden.py``; vectors committed under
  ``tests/golden/``) and against the known-answer vectors listed in SURVEY.md section 8c.
* ``ngp`` (the Instant-NGP glue, instant_ngp.py:129-263): PINNED to the reference's own
  ``InstantNGPPipeline``, constructed in a child process on the reference's ``HARP2Dataset`` around a
  stand-in ``tinycudann`` module that evaluates ``tcnn_spec`` (forward results, loss, parameter
  gradients, extract; ``tests/test_reference_interchange.py``, build container only).
* ``tcnn_spec`` (multiresolution hash grid, spherical harmonics, identity/composite
  encodings, bias-free fully fused MLP): **PARITY UNPINNED**.  That arithmetic lives in the
  third-party, un-vendored and un-pinned ``tiny-cuda-nn`` (reference README.md:19-22 installs
  git HEAD; ``import tinycudann`` at src/atmonr/pipelines/instant_ngp.py:4).  It is absent
  from /root/reference and cannot be installed here (no network), and the reference has no
  test or golden vector at that boundary.  ``tcnn_spec`` restates tiny-cuda-nn's published
  algorithm (Mueller et al. 2022 + upstream source as recalled; every item is marked
  "upstream-recalled") and is anchored on the reference's call sites
  (instant_ngp.py:60-85,163-174,236-237) and config (configs/instant_ngp.json:18-81).
"""
