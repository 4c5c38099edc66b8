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

PARITY UNPINNED for the e3nn / PyG parts: the reference ships no tests, golden
vectors or known-answer files for this path and e3nn/PyG cannot be imported
here, so those semantics are pinned only by (i) analytic properties (3j
invariance, D(R1 R2) = D(R1) D(R2), Y(Rx) = D(R) Y(x), SE(3) bi-equivariance of
the final scores) and (ii) the spot values recorded in SURVEY.md App. A.3.
The parts of the reference that DO import here (``diffusion_edf/transforms.py``
and ``diffusion_edf/radial_func.py``) are pinned by golden vectors generated
from the reference itself: ``tests/golden/make_golden.py``.
"""
