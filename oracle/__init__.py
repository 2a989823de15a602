"""CPU oracle for the spherical-projection hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy restatement of the reference algorithms
(hsientzucheng/CP-360-Weakly-Supervised-Saliency):

  oracle.cubepad   model/cube_pad.py:23-216            (CubePad index map + apply)
  oracle.e2c       utils/equi_to_cube.py:12-129        (Equi2Cube maps, cv2 fixed-point bilinear)
  oracle.c2e       utils/cube_to_equi.py:12-66         (Cube2Equi maps, single-pass grid_sample)
  oracle.cam       static_model/class_activation_model.py:46-52,76-90 (CAM contraction, heat-map post-ops;
                   pinned by tests/golden/make_golden_cam.py -> golden_cam.npz)
  oracle.ref_port  the same three ops re-expressed with the library calls the reference
                   itself makes (cv2.remap, torch.cat/flip, F.grid_sample) — CPU baseline
                   arm of bench.py only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package, and only as the checker or the timed CPU baseline.
The product package never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference ITSELF, executed in the build container from
/root/reference by ``tests/golden/make_golden.py``; the resulting fixtures live in
``tests/golden/*.npz|json`` and ``tests/test_oracle_golden.py`` checks the oracle against
every one of them (bit-exact for CubePad and the integer maps, <=1e-6 for float maps).
In the build container ``tests/test_oracle_vs_reference.py`` additionally runs the reference live on
seeded random configurations, ``tests/test_golden_reproducible.py`` regenerates the fixtures, and
``tests/golden/compare_port_vs_reference.py`` times ref_port next to the reference (same outputs, same speed).
"""
