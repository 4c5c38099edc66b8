"""Generate golden vectors from the REFERENCE itself (run in the authoring
container only: needs /root/reference).  Only ``diffusion_edf/transforms.py``
and ``diffusion_edf/radial_func.py`` import there (e3nn / PyG are absent), so
those are the functions pinned here; the output ``ref_golden.npz`` is committed
and is what the CPU tests load -- /root/reference is never read at test time.

    PYTHONPATH=/root/reference python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/diffusion_edf"


def _load(name):
    # import the two modules directly by path: the package __init__ is harmless but
    # sibling modules pull e3nn, so avoid the package import machinery altogether
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", os.path.join(REF, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    T = _load("transforms")
    R = _load("radial_func")
    g = torch.Generator().manual_seed(1234)
    out = {}
    q = torch.randn(64, 4, generator=g, dtype=torch.float64)
    p = torch.randn(64, 3, generator=g, dtype=torch.float64)
    q2 = torch.randn(64, 4, generator=g, dtype=torch.float64)
    qn = T.normalize_quaternion(q)
    out["q"], out["p"], out["q2"] = q, p, q2
    out["quaternion_to_matrix"] = T.quaternion_to_matrix(q)
    out["quaternion_raw_multiply"] = T.quaternion_raw_multiply(q, q2)
    out["quaternion_invert"] = T.quaternion_invert(q)
    out["quaternion_apply"] = T.quaternion_apply(qn, p)
    out["standardize_quaternion"] = T.standardize_quaternion(q)
    out["normalize_quaternion"] = qn
    out["euler_yxy"] = T.matrix_to_euler_angles(T.quaternion_to_matrix(qn), "YXY")

    x = torch.linspace(-0.5, 12.0, 501, dtype=torch.float64)
    out["x"] = x
    out["soft_step"] = R.soft_step(x / 10.0)
    out["ssc2_right"] = R.soft_square_cutoff_2(x, (None, None, 8.0, 10.0))
    out["ssc2_left"] = R.soft_square_cutoff_2(x, (0.06, 0.3, None, None))
    out["ssc_finite"] = R.soft_square_cutoff(x / 10.0, thr=0.8, infinite=False)
    out["ssc_infinite"] = R.soft_square_cutoff(x / 10.0, thr=0.8, infinite=True)

    xf = x.float().clamp(min=0)
    torch.manual_seed(0)
    grb = R.GaussianRadialBasis(dim=64, max_val=10.0).eval()
    out["gaussian_radial_basis_64_r10"] = grb(xf).detach()
    fin = R.GaussianRadialBasisLayerFiniteCutoff(num_basis=32, cutoff=0.99 * 3.0).eval()
    out["gaussian_finite_cutoff_32_r3"] = fin(xf[xf <= 3.0]).detach()
    out["xf_le3"] = xf[xf <= 3.0]
    sin = R.SinusoidalPositionEmbeddings(dim=64, max_val=100.0, n=1000.0)
    out["sinusoidal_64_100_1000"] = sin(xf * 8.0)
    sin_t = R.SinusoidalPositionEmbeddings(dim=256, max_val=1.0, n=10000.0)
    tt = torch.linspace(0.01, 1.0, 37, dtype=torch.float64)
    out["t"] = tt
    out["sinusoidal_256_1_10000_f64"] = sin_t(tt)
    out["xf"] = xf
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_golden.npz"),
                        **{k: v.numpy() for k, v in out.items()})
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
