"""CPU oracle for the Diffusion-EDF score-network hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``diffusion_edf_b200/`` may import this
package; it is imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the CPU baseline, never as the product path.

It restates, in plain PyTorch on the CPU (no e3nn / torch_scatter /
torch_cluster, which are absent from this image), the arithmetic of
``/root/reference/diffusion_edf`` for ``MultiscaleScoreModel.forward`` /
``ScoreModelBase.sample`` plus the semantics
of the un-vendored third-party ops the reference calls (SURVEY.md App. A:
e3nn==0.4.4 ``o3.TensorProduct`` / ``o3.SphericalHarmonics`` / ``wigner_3j`` /
``_Jd`` / ``normalize2mom``; torch_scatter ``scatter`` / ``scatter_logsumexp``;
torch_cluster ``radius`` / ``radius_graph`` / ``fps``).

What pins it:

* the reference's MODULE code (UNet / forward-only encoder, tensor field, attention
  blocks, score heads, keypoint extractor, point-attentive model, train loss,
  denoise loop) -- by the reference's own source, made executable with stand-ins
  for the absent third-party libraries: ``tests/golden/ref_shim.py`` +
  ``make_golden_model.py`` -> ``ref_model_golden.npz`` (all six shipped model
  families), held by ``tests/test_oracle.py::test_oracle_matches_reference_code_golden``;
* ``diffusion_edf/transforms.py`` and ``diffusion_edf/radial_func.py`` (they import
  as they are) -- golden vectors from the reference: ``tests/golden/make_golden.py``;
* ``voxel_filter`` -- the reference's function source on its own test scene:
  ``tests/golden/make_golden_voxel.py``.

PARITY UNPINNED for the arithmetic INSIDE e3nn / torch_scatter / torch_cluster:
the reference ships no tests or known-answer files for them and they cannot be
imported here, so those semantics (and the stand-ins above, which are built on
them) rest on (i) second sources present in the image -- scipy spherical
harmonics, sympy SU(2) Clebsch-Gordan and Wigner D, cKDTree radius search, scipy
logsumexp (``tests/test_oracle_independent.py``), (ii) analytic properties (3j
invariance, D(R1 R2) = D(R1) D(R2), Y(Rx) = D(R) Y(x), SE(3) bi-equivariance of
the final scores) and (iii) the spot values recorded in SURVEY.md App. A.3.
"""
