"""The CUDA path against the numbers the REFERENCE'S OWN SOURCE computes (tests/golden/ref_model_golden.npz, written by
tests/golden/make_golden_model.py from /root/reference with stand-ins for the un-installable third-party libraries).  Sorted
last on purpose: it was added after the round's GPU minutes were spent and has not run on a GPU yet; every step of it is the
composition of comparisons that have (CUDA vs oracle in test_gpu_model.py, oracle vs this fixture in test_oracle.py)."""
import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["pick", "place", "highres", "sapien_highres", "sapien_lowres", "ebm"])
def test_cuda_matches_reference_code_golden(cuda, kind):
    """The CUDA path against tests/golden/ref_model_golden.npz -- the numbers the REFERENCE'S OWN SOURCE computes for every shipped
    model family (tests/golden/make_golden_model.py; third-party ops supplied by the oracle's restatement): key scales, query
    points, scores (or critic energies), training loss + statistics, and the zero-temperature denoise trajectory."""
    import os
    import numpy as np
    import diffusion_edf_b200 as P
    from tests.golden.model_cases import SAMPLE_KW, feature_rows, inputs, seeded_oracle, spec, weight_checksums
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_model_golden.npz"))
    g = lambda k: torch.from_numpy(G[f"{kind}/{k}"])                      # noqa: E731
    kwargs, cls, has_scores, has_sample = spec(kind)
    oracle = seeded_oracle(kind)
    if not np.allclose(weight_checksums(oracle.state_dict()), G[f"{kind}/weights"], rtol=1e-9, atol=0):
        pytest.skip("this torch build draws different initial weights from seed 0 than the one the fixture was made with")
    model = getattr(P, cls)(**kwargs, deterministic=True).eval()
    model.load_state_dict(oracle.state_dict())
    model = model.to(cuda)
    model.requires_grad_(False)
    x, rgb, b, Ts, t, gx, gf, gb = (v.to(cuda) for v in inputs(kind))
    key, grasp = P.FeaturedPoints(x, rgb, b), P.FeaturedPoints(gx, gf, gb)
    with torch.no_grad():
        key_ms = model.get_key_pcd_multiscale(key)
        q = model.get_query_pcd(grasp)
        for s, p in enumerate(key_ms):
            assert torch.equal(p.x.cpu(), g(f"key{s}_x")), f"pooled coordinates of scale {s}"
            assert_close(p.f[feature_rows(len(p.x))], g(f"key{s}_f"), 1e-3, f"key features scale {s}")
            if f"{kind}/key{s}_w" in G.files:
                assert_close(p.w, g(f"key{s}_w"), 1e-3, f"key point weights scale {s}")
        assert torch.equal(q.x.cpu(), g("query_x"))
        assert_close(q.f, g("query_f"), 1e-3, "query features")
        assert_close(q.w, g("query_w"), 1e-3, "query weights")
        if has_scores:
            ang, lin = model.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=t)
            assert_close(ang, g("ang"), 1e-3, "ang")
            assert_close(lin, g("lin"), 1e-3, "lin")
            out = model.get_train_loss(Ts, t, key, grasp, g("target_ang").to(cuda), g("target_lin").to(cuda))
            loss, stats = out[0], out[-1]
            got = torch.tensor([float(loss)] + [float(stats[k]) for k in sorted(stats)], dtype=torch.float64)
            assert_close(got, g("loss"), 1e-3, "training loss and statistics")
        else:
            assert_close(model.score_head.compute_energy(Ts, key_ms, q, t), g("energy"), 1e-3, "critic energy")
        if has_sample:
            traj = model.sample(Ts, key_ms, q, **SAMPLE_KW)
            assert traj.shape == g("traj").shape and traj.dtype == torch.float64
            assert_close(traj.cpu(), g("traj"), 1e-3, "zero-temperature trajectory")
