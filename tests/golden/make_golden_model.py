"""Golden vectors from the REFERENCE'S OWN MODEL CODE (run in the authoring container only: needs /root/reference).

tests/golden/ref_shim.py makes the reference package importable by supplying stand-ins for the un-installable third-party
libraries (e3nn / torch_scatter / torch_cluster, implemented on the oracle's restatement of them).  This script then builds
the reference's `MultiscaleScoreModel` UNMODIFIED from /root/reference with the shipped panda_mug pick_lowres (and place_lowres)
kwargs, gives it the weights a seeded oracle model has (zero-initialised biases and unit layer-norm weights randomised), and records what the reference code computes:

  * `forward` (UNet encode + query model + score head): the four key scales (coordinates, features), the scores;
  * `get_train_loss`: the loss and its statistics;
  * `sample`: the zero-temperature (deterministic) pose trajectory over 2 x 3 steps.

tests/test_oracle.py::test_oracle_matches_reference_code_golden re-creates the oracle from the same seed and must
reproduce these numbers -- that pins the oracle's hand restatement of graph_parser / graph_attention / gnn_block /
multiscale_tensor_field / score_head / unet_feature_extractor / keypoint_extractor / score_model_base to the reference's source.
The weights are not stored (7 MB): the fixture carries checksums of them, and the test skips if a different torch RNG
stream produced different ones.

    python tests/golden/make_golden_model.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import ref_shim                                   # noqa: E402

ref_shim.install()

from diffusion_edf.gnn_data import FeaturedPoints as RefFP          # noqa: E402
from diffusion_edf.multiscale_score_model import MultiscaleScoreModel as RefModel    # noqa: E402
from diffusion_edf.point_attentive_score_model import PointAttentiveScoreModel as RefPointAttentive    # noqa: E402

from tests.golden.model_cases import KINDS, NO_LOSS, SAMPLE_KW, feature_rows, inputs, seeded_oracle, spec, weight_checksums    # noqa: E402

def run(kind, out):
    kwargs, cls, has_scores, has_sample = spec(kind)
    oracle = seeded_oracle(kind)
    sd = oracle.state_dict()
    ref = (RefPointAttentive if cls == "PointAttentiveScoreModel" else RefModel)(**kwargs, deterministic=True).eval()
    res = ref.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys[:5]
    # what the oracle does not carry are e3nn bookkeeping buffers only (output masks, empty weight buffers of weightless products)
    assert all(ref.state_dict()[k].numel() == 0 or k.endswith("output_mask") for k in res.missing_keys), \
        [k for k in res.missing_keys if ref.state_dict()[k].numel() and not k.endswith("output_mask")][:5]
    x, rgb, b, Ts, t, gx, gf, gb = inputs(kind)
    key, grasp = RefFP(x=x, f=rgb, b=b), RefFP(x=gx, f=gf, b=gb)
    out[f"{kind}/weights"] = weight_checksums(sd)
    with torch.no_grad():
        key_ms = ref.get_key_pcd_multiscale(key)
        q = ref.get_query_pcd(grasp)
        for s, p in enumerate(key_ms):
            out[f"{kind}/key{s}_x"], out[f"{kind}/key{s}_f"] = p.x.numpy(), p.f[feature_rows(len(p.x))].numpy()
            if p.w is not None:
                out[f"{kind}/key{s}_w"] = p.w.numpy()
        out[f"{kind}/query_x"], out[f"{kind}/query_f"], out[f"{kind}/query_w"] = q.x.numpy(), q.f.numpy(), q.w.numpy()
        msg = ""
        if has_scores:
            ang, lin = ref.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=t)
            out[f"{kind}/ang"], out[f"{kind}/lin"] = ang.numpy(), lin.numpy()
            msg = f"ang {ang.abs().max().item():.4f}"
            if kind not in NO_LOSS:
                g = torch.Generator().manual_seed(11)
                ta, tl = torch.randn(len(Ts), 3, generator=g), torch.randn(len(Ts), 3, generator=g)
                loss, _, _, stats = ref.get_train_loss(Ts, t, key, grasp, ta, tl)
                out[f"{kind}/target_ang"], out[f"{kind}/target_lin"] = ta.numpy(), tl.numpy()
                out[f"{kind}/loss"] = np.array([float(loss)] + [float(stats[k]) for k in sorted(stats)], dtype=np.float64)
                msg += f" loss {float(loss):.4f}"
        else:
            e = ref.score_head.compute_energy(Ts, key_ms, q, t)
            out[f"{kind}/energy"] = e.numpy()
            msg = f"energy {e.min().item():.4f}..{e.max().item():.4f}"
            # EbmScoreModelHead.forward (score_head_ebm.py:192-222): the score as the pose gradient of -energy (autograd through
            # the reference's own field code; .eval() put the head in inference mode: first order, detached)
            with torch.enable_grad():
                ang, lin = ref.score_head(Ts=Ts, key_pcd_multiscale=key_ms, query_pcd=q, time=t)
            out[f"{kind}/ang"], out[f"{kind}/lin"] = ang.detach().numpy(), lin.detach().numpy()
            msg += f" ang {ang.abs().max().item():.4f} lin {lin.abs().max().item():.4f}"
        if has_sample:
            traj = ref.sample(Ts, scene_pcd_multiscale=key_ms, grasp_pcd=q, **SAMPLE_KW)
            out[f"{kind}/traj"] = traj.numpy()
    print(kind, msg, "key scales", [len(p.x) for p in key_ms], "query", len(q.x), "missing (bookkeeping only):", len(res.missing_keys))


def run_c1(out):
    """BASELINE.json configs[0]: the reference's MultiscaleTensorField on the lmax = 1 plumbing case."""
    from diffusion_edf.multiscale_tensor_field import MultiscaleTensorField as RefField
    from tests.golden.model_cases import C1_KWARGS, c1_inputs, c1_seeded_oracle
    oracle = c1_seeded_oracle()
    ref = RefField(**C1_KWARGS).eval()
    res = ref.load_state_dict(oracle.state_dict(), strict=False)
    assert not res.unexpected_keys and all(ref.state_dict()[k].numel() == 0 or k.endswith("output_mask") for k in res.missing_keys)
    x0, f0, xq = c1_inputs()
    z = lambda n: torch.zeros(n, dtype=torch.long)                   # noqa: E731
    keys = [RefFP(x=x0, f=f0, b=z(256)), RefFP(x=x0[:32], f=f0[:32], b=z(32))]
    with torch.no_grad():
        o = ref(query_points=RefFP(x=xq, f=torch.empty(64, 0), b=z(64)), input_points_multiscale=keys)
    out["c1/weights"] = weight_checksums(oracle.state_dict())
    out["c1/out_f"] = o.f.numpy()
    print("c1 out", tuple(o.f.shape), float(o.f.abs().max()))


def main():
    out = {}
    only = [a for a in sys.argv[1:] if a in KINDS]                  # `make_golden_model.py ebm`: refresh these kinds only
    if only:
        out.update(np.load(os.path.join(HERE, "ref_model_golden.npz")))
    else:
        run_c1(out)
    for kind in only or KINDS:
        run(kind, out)
    np.savez_compressed(os.path.join(HERE, "ref_model_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_model_golden.npz"), os.path.getsize(os.path.join(HERE, "ref_model_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
