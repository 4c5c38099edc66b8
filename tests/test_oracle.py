"""CPU tests of the oracle: golden vectors produced from the reference's own importable modules
(tests/golden/make_golden.py), SURVEY.md App. A.3 spot values, and the analytic properties that pin the
e3nn semantics (3j invariance, Wigner-D homomorphism, SH equivariance, SE(3) equivariance of the scores)."""
import copy
import math
import os

import numpy as np
import pytest
import torch

from oracle import encoders as enc
from oracle import graph as OG
from oracle import model as OM
from oracle import nn as ON
from oracle import so3
from oracle.irreps import Irreps, sort_even_first

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))


def g(name):
    return torch.from_numpy(GOLD[name])


# ------------------------------------------------------------------ golden vectors from the reference
def test_quaternion_helpers_match_reference():
    q, p, q2 = g("q"), g("p"), g("q2")
    assert torch.allclose(enc.quaternion_to_matrix(q), g("quaternion_to_matrix"), atol=1e-14)
    assert torch.allclose(enc.quaternion_raw_multiply(q, q2), g("quaternion_raw_multiply"), atol=1e-14)
    assert torch.equal(enc.quaternion_invert(q), g("quaternion_invert"))
    qn = enc.normalize_quaternion(q)
    assert torch.allclose(qn, g("normalize_quaternion"), atol=1e-15)
    assert torch.allclose(enc.quaternion_apply(qn, p), g("quaternion_apply"), atol=1e-14)
    assert torch.equal(enc.standardize_quaternion(q), g("standardize_quaternion"))
    assert torch.allclose(enc.matrix_to_euler_yxy(enc.quaternion_to_matrix(qn)), g("euler_yxy"), atol=1e-13)


def test_cutoffs_match_reference():
    x = g("x")
    assert torch.allclose(enc.soft_step(x / 10.0).double(), g("soft_step").double(), atol=1e-14)
    assert torch.allclose(enc.soft_square_cutoff_2(x, (None, None, 8.0, 10.0)).double(), g("ssc2_right").double(), atol=1e-14)
    assert torch.allclose(enc.soft_square_cutoff_2(x, (0.06, 0.3, None, None)).double(), g("ssc2_left").double(), atol=1e-14)
    assert torch.allclose(enc.soft_square_cutoff(x / 10.0, thr=0.8, infinite=False).double(), g("ssc_finite").double(), atol=1e-14)
    assert torch.allclose(enc.soft_square_cutoff(x / 10.0, thr=0.8, infinite=True).double(), g("ssc_infinite").double(), atol=1e-14)


def test_length_and_time_encoders_match_reference():
    xf = g("xf")
    grb = enc.GaussianRadialBasis(dim=64, max_val=10.0)
    assert torch.allclose(grb(xf), g("gaussian_radial_basis_64_r10"), atol=1e-6, rtol=1e-6)
    fin = enc.GaussianRadialBasisLayerFiniteCutoff(num_basis=32, cutoff=0.99 * 3.0)
    assert torch.allclose(fin(g("xf_le3")), g("gaussian_finite_cutoff_32_r3"), atol=1e-6, rtol=1e-6)
    sin = enc.SinusoidalPositionEmbeddings(dim=64, max_val=100.0, n=1000.0)
    assert torch.allclose(sin(xf * 8.0), g("sinusoidal_64_100_1000"), atol=1e-6)
    sin_t = enc.SinusoidalPositionEmbeddings(dim=256, max_val=1.0, n=10000.0)
    assert torch.allclose(sin_t(g("t")), g("sinusoidal_256_1_10000_f64"), atol=1e-12)


# ------------------------------------------------------------------ e3nn semantics (App. A)
def test_w3j_spot_values_and_nnz():
    w = so3.wigner_3j
    assert abs(w(1, 1, 1)[0, 1, 2].item() - 1 / math.sqrt(6)) < 1e-12          # eps_ijk / sqrt6, + for xyz cyclic
    assert torch.allclose(torch.diag(w(1, 1, 2)[:, :, 2]), torch.tensor([-0.18257419, 0.36514837, -0.18257419], dtype=torch.float64), atol=1e-7)
    assert abs(w(1, 1, 2)[0, 2, 0].item() - 0.31622777) < 1e-7 and abs(w(1, 1, 2)[2, 0, 0].item() - 0.31622777) < 1e-7
    assert abs(w(2, 2, 2)[2, 2, 2].item() - 0.23904572) < 1e-7 and abs(w(2, 2, 2)[0, 0, 2].item() + 0.23904572) < 1e-7
    assert abs(w(1, 2, 2)[0, 0, 1].item() + 0.18257419) < 1e-7 and abs(w(1, 2, 2)[0, 1, 0].item() - 0.18257419) < 1e-7
    assert abs(w(1, 2, 2)[0, 2, 3].item() - 0.31622777) < 1e-7 and abs(w(1, 2, 2)[0, 3, 4].item() - 0.18257419) < 1e-7
    nnz = {(0, 0, 0): 1, (0, 1, 1): 3, (0, 2, 2): 5, (1, 1, 0): 3, (1, 1, 1): 6, (1, 1, 2): 11, (1, 2, 2): 16, (2, 2, 0): 5,
           (2, 2, 1): 16, (2, 2, 2): 25}
    for ls, n in nnz.items():
        assert int((w(*ls).abs() > 1e-9).sum()) == n, ls
    for l in (1, 2):    # what e3nn 0.4.4's specialised code paths assume
        eye = torch.eye(2 * l + 1, dtype=torch.float64) / math.sqrt(2 * l + 1)
        assert torch.allclose(w(0, l, l)[0], eye, atol=1e-12) and torch.allclose(w(l, l, 0)[:, :, 0], eye, atol=1e-12)


def _rand_rot(n, seed):
    gen = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen, dtype=torch.float64), dim=-1)
    return q, enc.quaternion_to_matrix(q)


def test_wigner_d_properties():
    q, R = _rand_rot(8, 0)
    ang = enc.matrix_to_euler_yxy(R)
    for l in (1, 2):
        De = so3.wigner_D_euler(l, ang[:, 0], ang[:, 1], ang[:, 2])        # the reference's route (J matrices)
        Dm = so3.wigner_D_from_matrix(l, R)                                 # definition Y(Rx) = D Y(x)
        assert (De - Dm).abs().max() < 1e-12
        assert (Dm @ Dm.transpose(-1, -2) - torch.eye(2 * l + 1, dtype=torch.float64)).abs().max() < 1e-12
        D12 = so3.wigner_D_from_matrix(l, R[:4] @ R[4:])
        assert (D12 - Dm[:4] @ Dm[4:]).abs().max() < 1e-12                 # homomorphism
        J = so3.J_matrix(l)
        assert (J @ J - torch.eye(2 * l + 1, dtype=torch.float64)).abs().max() < 1e-12 and (J - J.T).abs().max() < 1e-12
    assert (so3.wigner_D_from_matrix(1, R) - R).abs().max() < 1e-12        # D^1 = R in (x,y,z) order (score_head.py:198-205 relies on it)


def test_w3j_invariance():
    _, R = _rand_rot(4, 1)
    for ls in [(1, 1, 1), (1, 1, 2), (1, 2, 1), (1, 2, 2), (2, 1, 1), (2, 1, 2), (2, 2, 0), (2, 2, 1), (2, 2, 2)]:
        C = so3.wigner_3j(*ls)
        D = [so3.wigner_D_from_matrix(l, R) for l in ls]
        C2 = torch.einsum("tia,tjb,tkc,abc->tijk", D[0], D[1], D[2], C)
        assert (C2 - C).abs().max() < 1e-12, ls


def test_normalize2mom_constants():
    c = ON.act_consts()
    assert abs(c["silu"] - 1.6791767923989418) < 1e-12
    assert abs(c["sigmoid"] - 1.8467055342154763) < 1e-12
    assert abs(c["slrelu"] - 1.531320475574866) < 1e-12


def test_depthwise_tp_layout_matches_survey_appendix_e():
    irr = Irreps("64x0e+32x1e+16x2e")
    dtp = ON.DepthwiseTensorProduct(irr, Irreps("1x0e+1x1e+1x2e"), irr, internal_weights=False, bias=False)
    assert str(dtp.irreps_out.simplify()) == "112x0e+192x1e+176x2e" and dtp.tp.weight_numel == 480
    i_out = [io for (_, _, io, _) in dtp.tp.instructions]
    assert i_out == [0, 3, 9, 4, 1, 5, 10, 6, 11, 12, 7, 13, 2, 8, 14]
    offs = [s.start for s in dtp.irreps_out.slices()]
    assert [offs[i] for i in i_out] == [0, 112, 688, 304, 64, 400, 1008, 496, 1168, 1328, 592, 1408, 96, 640, 1488]
    head = ON.DepthwiseTensorProduct(irr, irr, Irreps("1x0e+32x1e"), internal_weights=True, bias=False)
    assert head.tp.weight_numel == 11776 and str(head.irreps_out.simplify()) == "112x0e+192x1e"


def test_scatter_and_graph_ops():
    src = torch.tensor([[1.0, -2.0], [3.0, 0.5], [0.0, 0.0]])
    idx = torch.tensor([2, 2, 0])
    lse = OG.scatter_logsumexp(src, idx, 4)
    assert torch.allclose(lse[2], torch.logsumexp(src[:2], 0), atol=1e-6) and torch.equal(lse[1], torch.zeros(2)) and torch.equal(lse[3], torch.zeros(2))
    x = torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [0, 2.0, 0], [5.0, 5, 5]])
    e = OG.radius(x, x[:2], 1.5, None, None, 10)
    assert e.tolist() == [[0, 0, 1, 1], [0, 1, 0, 1]]
    assert OG.radius_graph(x, 1.5, None, False, 10).tolist() == [[0, 1], [1, 0]]
    assert OG.radius(x, x[:1], 100.0, None, None, 2).tolist() == [[0, 0], [0, 1]]      # truncation keeps the first in index order
    assert OG.fps(x, None, 0.5).tolist() == [0, 3]


# ------------------------------------------------------------------ model-level properties
def _small_model(seed=0):
    from diffusion_edf_b200.synthetic import model_kwargs
    torch.manual_seed(seed)
    return OM.MultiscaleScoreModel(**model_kwargs(), deterministic=True).eval()


def test_config_c1_tensor_field_plumbing():
    """BASELINE config C1: MultiscaleTensorField 1 layer, lmax=1, 256-point cloud, CPU."""
    torch.manual_seed(0)
    tf = OM.MultiscaleTensorField(irreps_input="16x0e+8x1e", irreps_output="16x0e+8x1e", irreps_sh="1x0e+1x1e", num_heads=4,
                                  fc_neurons=[-1, 16, 16], length_emb_dim=16, irreps_query=None, edge_context_emb_dim=None,
                                  r_cluster_multiscale=[2.0, None], length_enc_max_r=10.0, r_mincut_nonscalar_sh=0.1, n_layers=1).eval()
    x0 = torch.rand(256, 3) * 6 - 3
    f0 = torch.randn(256, 40)
    keys = [OM.FeaturedPoints(x0, f0, torch.zeros(256, dtype=torch.long)), OM.FeaturedPoints(x0[:32], f0[:32], torch.zeros(32, dtype=torch.long))]
    xq = torch.rand(64, 3) * 6 - 3
    q = OM.FeaturedPoints(xq, torch.empty(64, 0), torch.zeros(64, dtype=torch.long))
    with torch.no_grad():
        out = tf(q, keys)
        assert out.f.shape == (64, 40) and torch.isfinite(out.f).all()
        # rotate everything: scalars invariant, vectors rotate
        _, R = _rand_rot(1, 5)
        R = R[0].float()
        D = torch.block_diag(torch.eye(16), *([R] * 8))
        keys_r = [OM.FeaturedPoints(k.x @ R.T, k.f @ D.T, k.b) for k in keys]
        out_r = tf(OM.FeaturedPoints(xq @ R.T, q.f, q.b), keys_r)
    assert (out_r.f - out.f @ D.T).abs().max() < 2e-5 * out.f.abs().max().clamp_min(1)


def test_score_model_se3_equivariance_and_shapes():
    from diffusion_edf_b200.synthetic import make_poses, make_scene
    m = _small_model()
    assert sum(p.numel() for p in m.parameters()) == 1839556
    x, rgb = make_scene(1200, seed=1, half_extent=10.0)
    Ts, t = make_poses(3, x, seed=1, spread=4.0)
    b = torch.zeros(len(x), dtype=torch.long)
    grasp = OM.FeaturedPoints(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, dtype=torch.long))
    with torch.no_grad():
        (ang, lin), dbg = m(Ts, t, OM.FeaturedPoints(x, rgb, b), grasp, debug=True)
        assert ang.shape == lin.shape == (3, 3)
        assert [len(p.x) for p in dbg[0]] == [240, 48, 10, 2] and all(p.f.shape[1] == 240 for p in dbg[0])
        gq, R = _rand_rot(1, 2)
        gq, R = gq.float(), R[0].float()
        tg = torch.tensor([3.0, -2.0, 1.0])
        Ts2 = torch.cat([enc.quaternion_raw_multiply(gq.expand(3, -1), Ts[:, :4]), Ts[:, 4:] @ R.T + tg], -1)
        (ang2, lin2), _ = m(Ts2, t, OM.FeaturedPoints(x @ R.T + tg, rgb, b), grasp)
    # scores are expressed in the body frame -> invariant under a left action on (scene, poses)
    assert (ang2 - ang).abs().max() < 1e-4 * ang.abs().max() + 1e-6
    assert (lin2 - lin).abs().max() < 1e-4 * lin.abs().max() + 1e-6


def test_sample_zero_temperature_is_deterministic_and_shaped():
    from diffusion_edf_b200.synthetic import make_poses, make_scene
    m = _small_model(1)
    x, rgb = make_scene(600, seed=2, half_extent=7.0)
    T0, _ = make_poses(2, x, seed=2, spread=3.0)
    b = torch.zeros(len(x), dtype=torch.long)
    with torch.no_grad():
        keys = m.get_key_pcd_multiscale(OM.FeaturedPoints(x, rgb, b))
        q = m.get_query_pcd(OM.FeaturedPoints(torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, dtype=torch.long)))
        kw = dict(diffusion_schedules=[[1.0, 0.5]], N_steps=[3], timesteps=[0.04], temperatures=[0.0])
        a = m.sample(T0, keys, q, **kw)
        b2 = m.sample(T0, keys, q, **kw)
    assert a.shape == (5, 2, 7) and a.dtype == torch.float64 and torch.equal(a, b2) and torch.equal(a[-1], a[-2])
    assert (a[1:, :, :4].norm(dim=-1) - 1).abs().max() < 1e-12      # row 0 is the fp32-normalised seed itself


def test_voxel_filter_matches_reference_golden():
    """oracle.graph.voxel_filter against vectors produced by the reference's own voxel_filter source on the reference's test
    scene (tests/golden/make_golden_voxel.py): bit-exact, both coordinate reductions."""
    import numpy as np
    from oracle.graph import voxel_filter
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "voxel_golden.npz"))
    p, c = torch.tensor(g["points"]), torch.tensor(g["colors"])
    for red in ("average", "center"):
        co, fe = voxel_filter(p, c, float(g["voxel_size"]), red)
        assert torch.equal(co, torch.tensor(g[f"coord_{red}"])) and torch.equal(fe, torch.tensor(g[f"feat_{red}"]))


# ------------------------------------------------------------------ the reference's own model code (tests/golden/make_golden_model.py)
@pytest.mark.parametrize("kind", ["pick", "place", "highres", "sapien_highres", "sapien_lowres", "ebm", "pick_c2", "pick_1024"])
def test_oracle_matches_reference_code_golden(kind):
    """ref_model_golden.npz holds what the REFERENCE'S OWN SOURCE computes (key scales, query points, scores or energies,
    get_train_loss, zero-temperature sample) for every shipped model family when its un-installable third-party libraries are
    replaced by stand-ins built on oracle/so3.py and oracle/graph.py (tests/golden/ref_shim.py).  The oracle, re-created from
    the same seed, must reproduce it: this pins the hand restatement of the reference's module code (UNet / forward-only
    encoder, tensor field, attention blocks, score heads, keypoint extractor, point-attentive model, denoise loop)."""
    from tests.golden.model_cases import NO_LOSS, SAMPLE_KW, feature_rows, inputs, seeded_oracle, spec, weight_checksums
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_model_golden.npz"))
    g = lambda k: torch.from_numpy(G[f"{kind}/{k}"])                      # noqa: E731
    _, _, has_scores, has_sample = spec(kind)
    oracle = seeded_oracle(kind)
    if not np.allclose(weight_checksums(oracle.state_dict()), G[f"{kind}/weights"], rtol=1e-9, atol=0):
        pytest.skip("this torch build draws different initial weights from seed 0 than the one the fixture was made with")
    x, rgb, b, Ts, t, gx, gf, gb = inputs(kind)
    key, grasp = OM.FeaturedPoints(x, rgb, b), OM.FeaturedPoints(gx, gf, gb)

    def close(a, ref, tol, what):
        err = float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        assert err < tol, f"{kind}/{what}: {err:.3e}"

    with torch.no_grad():
        key_ms = oracle.get_key_pcd_multiscale(key)
        q = oracle.get_query_pcd(grasp)
        assert len(key_ms) == sum(1 for k in G.files if k.startswith(f"{kind}/key") and k.endswith("_x"))
        for s, p in enumerate(key_ms):
            assert torch.equal(p.x, g(f"key{s}_x")), f"pooled coordinates of scale {s}"
            close(p.f[feature_rows(len(p.x))], g(f"key{s}_f"), 2e-5, f"key features scale {s}")
            if f"{kind}/key{s}_w" in G.files:
                close(p.w, g(f"key{s}_w"), 2e-5, f"key point weights scale {s}")
        assert torch.equal(q.x, g("query_x"))
        close(q.f, g("query_f"), 2e-5, "query features")
        close(q.w, g("query_w"), 2e-5, "query weights")
        if has_scores:
            ang, lin = oracle.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=t)
            close(ang, g("ang"), 2e-5, "ang")
            close(lin, g("lin"), 2e-5, "lin")
            if kind not in NO_LOSS:
                out = oracle.get_train_loss(Ts, t, key, grasp, g("target_ang"), g("target_lin"))
                close(torch.tensor([float(out[0])], dtype=torch.float64), g("loss")[:1], 2e-5, "training loss")     # g("loss")[1:]: the statistics dict
        else:
            close(oracle.score_head.compute_energy(Ts, key_ms, q, t), g("energy"), 2e-5, "energy")
            ang, lin = oracle.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=t)      # EbmScoreModelHead.forward
            close(ang, g("ang"), 5e-5, "ebm ang")
            close(lin, g("lin"), 5e-5, "ebm lin")
        if has_sample:
            traj = oracle.sample(Ts, key_ms, q, **SAMPLE_KW)
            assert traj.shape == g("traj").shape and traj.dtype == torch.float64
            close(traj, g("traj"), 1e-4, "zero-temperature trajectory")       # fp32 scores integrated over 6 steps (observed 1e-5)


def test_config_c1_matches_reference_code_golden():
    """BASELINE.json configs[0] (lmax = 1 MultiscaleTensorField, 256-point cloud, CPU): the oracle against the output of the
    reference's own MultiscaleTensorField source (tests/golden/make_golden_model.py::run_c1)."""
    from tests.golden.model_cases import c1_inputs, c1_seeded_oracle, weight_checksums
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_model_golden.npz"))
    tf = c1_seeded_oracle()
    if not np.allclose(weight_checksums(tf.state_dict()), G["c1/weights"], rtol=1e-9, atol=0):
        pytest.skip("this torch build draws different initial weights from seed 0 than the one the fixture was made with")
    x0, f0, xq = c1_inputs()
    z = lambda n: torch.zeros(n, dtype=torch.long)                   # noqa: E731
    keys = [OM.FeaturedPoints(x0, f0, z(256)), OM.FeaturedPoints(x0[:32], f0[:32], z(32))]
    with torch.no_grad():
        out = tf(OM.FeaturedPoints(xq, torch.empty(64, 0), z(64)), keys)
    ref = torch.from_numpy(G["c1/out_f"])
    assert out.f.shape == ref.shape == (64, 40)
    assert float((out.f - ref).abs().max() / ref.abs().max()) < 2e-5


# ------------------------------------------------------------------ collision-aware trajectory optimisation
@pytest.mark.parametrize("name", ["knn", "radius", "knn_sparse"])
def test_collision_oracle_matches_reference_code_golden(name):
    """collision_golden.npz holds what the REFERENCE'S OWN function sources compute (collision_utils._pcd_energy,
    _optimize_pcd_collision_trajectory, _check_pcd_collision with se3._exp_map / _multiply and pcd_utils.transform_points, executed by
    tests/golden/make_golden_collision.py with stand-ins for torch_cluster.knn / radius and torch_scatter.scatter_sum only);
    oracle/collision.py must reproduce it."""
    from oracle import collision as OC
    from tests.golden.collision_cases import CASES, inputs
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "collision_golden.npz"))
    cfg = CASES[name]
    x, y, Ts = inputs(name)
    Ty = OC.transform_points_batched(y.expand(len(Ts), -1, 3), Ts)
    e, g = OC.pcd_energy(x, Ty, cfg["cutoff_r"], cfg["k"], cfg["eps"], True, cfg["method"])
    tr = OC.optimize_trajectory(x, y, Ts, cfg["n_steps"], cfg["dt"], cfg["cutoff_r"], cfg["k"], cfg["eps"], cfg["method"], revert_order=True)
    hit = OC.check_pcd_collision(x, Ty, cfg["check_r"])
    rel = lambda a, k: float((a - torch.from_numpy(G[f"{name}/{k}"])).abs().max() / max(1e-30, float(np.abs(G[f"{name}/{k}"]).max())))   # noqa: E731
    assert rel(e, "energy") < 1e-6 and rel(g, "grad") < 1e-6 and rel(tr, "traj") < 1e-6
    assert np.array_equal(hit.numpy(), G[f"{name}/hit"])
    assert float(e[-1]) == 0.0 and torch.equal(tr[-1, 0], tr[-1, -1])          # the far-away pose: no energy, never moves
    assert float(e[:-1].min()) > 0.0


def test_collision_energy_gradient_is_the_lie_derivative():
    """Independent pin of _pcd_energy's gradient convention: a finite difference of the energy under a small world-frame rotation /
    translation of the transformed cloud (float64, radius method with a neighbour set that does not change)."""
    from oracle import collision as OC
    g = torch.Generator().manual_seed(5)
    x = torch.rand(400, 3, generator=g, dtype=torch.float64) * 2 - 1
    y = (torch.rand(2, 30, 3, generator=g, dtype=torch.float64) * 2 - 1) * 0.5
    e0, grad = OC.pcd_energy(x, y, 0.5, max_num_neighbor=10_000, eps=0.05, cluster_method="radius")
    h = 1e-6
    for a in range(3):
        d = torch.zeros(3, dtype=torch.float64); d[a] = h
        e_t, _ = OC.pcd_energy(x, y + d, 0.5, max_num_neighbor=10_000, eps=0.05, compute_grad=False, cluster_method="radius")
        e_r, _ = OC.pcd_energy(x, y + torch.cross(d.expand_as(y), y, dim=-1), 0.5, max_num_neighbor=10_000, eps=0.05, compute_grad=False,
                               cluster_method="radius")
        assert torch.allclose((e_t - e0) / h, grad[:, 3 + a], rtol=2e-3, atol=1e-3 * float(grad.abs().max()))
        assert torch.allclose((e_r - e0) / h, grad[:, a], rtol=2e-3, atol=1e-3 * float(grad.abs().max()))
