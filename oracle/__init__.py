"""CPU oracle for the fs-nerf ray-march hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain fp32 CPU restatement (numpy + torch-CPU) of the
algorithms on the hot path named in BASELINE.json.  It is the *checker* for the
CUDA kernels in ``fsnerf_b200/csrc`` — it is never imported by the product
package ``fsnerf_b200``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Pinning status (see DESIGN.md "Oracle"):

* ``rays``, ``encoding``, ``mlp``  — PINNED against the reference's own Python
  (``/root/reference/src/utils/utilities.py``, ``src/core/models.py``) imported
  in the build container by ``oracle/gen_golden.py``; outputs committed under
  ``tests/golden/reference_*.npz`` and compared bit-for-bit / to 1e-6.
* ``compositing.render_packed`` — restates ``nerfacc.volrend.rendering`` v0.5.3
  (third-party, ``environment.yaml:341``, source NOT on the box).  Pinned only at
  the reference's call site (``src/render/rendering.py:89-96``) and by
  closed-form known answers.  **parity unpinned** against nerfacc itself.
* ``sampling`` (stratified, sample_pdf), ``encoding.freq_mask`` — no reference
  counterpart exists; canonical NeRF / FreeNeRF
  definitions are the spec.  **parity unpinned.**
"""
